"""TEST / BASELINE INFRASTRUCTURE ONLY.  Make the UNMODIFIED reference available on the GPU box.

``/root/reference`` only exists in the build container.  ``__graft_entry__.build()`` calls ``fetch()`` here, which
copies the reference's Python sources (nothing else: no logs, no images) verbatim into ``baseline/_ref/`` -- a
git-IGNORED directory (the history stays free of reference sources) that is not gpurun-ignored, so it travels with
the repo snapshot to the B200 box.  There it serves three purposes, none of them on the product path:

  * ``bench.py --impl reference``: the reference's own ``dis_update`` + ``gen_update`` timed on the host cores
    (``cpu_baseline.kind == "reference"``) and, informatively, on the B200 through its stock cuDNN path;
  * ``-m gpu`` parity tests that compare full tensors with the live reference at the benchmarked sizes;
  * the drop-in test that execs the unmodified ``main.py`` on top of the product's modules.

Everything that uses it degrades to the committed ``tests/golden`` fixtures / the ``oracle/restate.py`` port when
the directory is absent.
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("src_deformable", "src_baseline")


def fetch(verbose=False):
    """Copy every ``*.py`` of the two reference trees (plus README / commands for context) into baseline/_ref."""
    if not os.path.isdir(os.path.join(SRC, "src_deformable")):
        return False
    n = 0
    for tree in TREES:
        for dirpath, dirnames, filenames in os.walk(os.path.join(SRC, tree)):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "logs", "tests")]
            rel = os.path.relpath(dirpath, SRC)
            for f in filenames:
                if not (f.endswith(".py") or f == "commands"):
                    continue
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                dst = os.path.join(DST, rel, f)
                src = os.path.join(dirpath, f)
                if not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or \
                        os.path.getsize(dst) != os.path.getsize(src):
                    shutil.copyfile(src, dst)
                n += 1
    if verbose:
        print("baseline/_ref: %d reference source files" % n)
    return True


def root(tree="src_deformable"):
    """Directory of a reference tree: the mounted reference if present, else the travelling copy, else None."""
    for base in (SRC, DST):
        p = os.path.join(base, tree)
        if os.path.isfile(os.path.join(p, "models", "networks.py")):
            return p
    return None


if __name__ == "__main__":
    fetch(verbose=True)
