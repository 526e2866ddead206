"""TEST INFRASTRUCTURE ONLY — CPU oracle for the deformable-GAN training step.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker
or as the timed CPU baseline.  The product path (``pose_transfer_b200``) never imports this package.

Modules
-------
ref_import : shim that imports the *unmodified* reference (``/root/reference/src_deformable``) in the
             build container; used to generate ``tests/golden`` and to pin ``restate``.
synth      : seeded synthetic batches / weights shared by the oracle, the tests and bench.py.
restate    : pure-torch fp32 CPU restatement of the reference algorithm (file:line cited per function).
"""
