"""pose_transfer_b200 -- B200 (sm_100a) implementation of the deformable-GAN training step of
saurabhsharma1993/pose-transfer (src_deformable, warp_skip=mask), behind the reference's own module
surface (models.networks / models.pose_gan / utils.pose_transform / utils.pose_utils).

Python owns tensors, module/state_dict structure, the optimiser object and the NCCL process group;
every FLOP and byte on the step goes through hand-written CUDA kernels in ``csrc/libptk.so`` (C ABI in
``include/ptk.h``).  There is no CPU or eager-PyTorch fallback: calling a kernel without the built library
or without a CUDA device raises.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the .so until first use)
