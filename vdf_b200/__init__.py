"""vdf_b200 -- B200-native (sm_100a) data-parallel hot path of protocol/vdf's Nova prover for MinRoot.

Host-side mirror of the reference's interfaces for this path (src/minroot.rs, src/nova/proof.rs) over the
C ABI of libvdfgpu.so (include/vdfgpu.h).  The GPU path has no CPU fallback.
"""
from . import encoding  # noqa: F401
from ._lib import VdfGpuError, load  # noqa: F401

__all__ = ["encoding", "load", "VdfGpuError", "minroot", "msm", "nova"]
