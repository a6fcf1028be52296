"""pseudospectra.jl_b200 -- B200-native resolvent-norm grid kernel behind Pseudospectra.jl's psacore/psa_compute
interface.  The directory name contains a dot, so import it through `psa_b200_loader.load_package()` at the
repository root (tests, bench.py and __graft_entry__.py all do)."""
from ._lib import (PsaGpu, PsaGpuCancelled, PsaGpuError, PsaGpuUnsupported, load as load_library, version,  # noqa: F401
                   EXPORTS, LIB_PATH)
from .compute import ComputeThresholds, psa_compute, psacore, psmode_inv_lanczos, vec2ax  # noqa: F401
