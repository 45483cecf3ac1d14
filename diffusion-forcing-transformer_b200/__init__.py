"""dfot_b200 — B200-native (sm_100a) implementation of the DFoT denoising sampling path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic on the
hot path runs in hand-written CUDA kernels reached through the C ABI in ``include/dfot_b200.h``.
There is no CPU fallback: using a kernel without the built library / a CUDA device raises.

Module layout mirrors the reference (``algorithms.dfot...``) so that
``from dfot_b200.algorithms.dfot import DFoTVideo`` is the drop-in for
``from algorithms.dfot import DFoTVideo``.
"""
__version__ = "0.1.0"
