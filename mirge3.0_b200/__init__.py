"""mirge3.0_b200 -- B200-native (sm_100a) implementation of miRge3.0's per-read hot path:
digest (FASTQ parse + cutadapt-semantics trimming) -> collapse -> ordered annotation rounds,
behind miRge3.0's own entry points ``baking`` (mirge/libs/digest.py:105) and ``bwtAlign``
(mirge/libs/manifoldAlign.py:68).  The arithmetic runs in hand-written CUDA kernels reached through
the C ABI declared in ``include/mirge_b200.h``; PyTorch only provides device buffers and streams.
There is no CPU fallback: importing the compute modules without the built library raises."""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PACKAGE_DIR)
LIB_PATH = os.path.join(PACKAGE_DIR, "libmirge_b200.so")

__all__ = ["PACKAGE_DIR", "REPO_ROOT", "LIB_PATH"]
