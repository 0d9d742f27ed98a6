"""depecher_b200 — B200-native (sm_100a CUDA) engine for DepecheR's penalized k-means hot path.

Only what the path needs lives here:
  csrc/        hand-written CUDA kernels + the C ABI (include/depecher_b200.h) -> lib/libdepecher_b200.so
  _capi.py     ctypes binding of that ABI (what an Rcpp layer binds)
  rapi.py      host-side mirror of the reference's R callers of the path: depeche(), dAllocate(),
               depechePenaltyOpt(), depecheAllData(), depecheClusterCenterSelection()
"""
from ._capi import ARI_EXACT, ARI_MONTE_CARLO, Dataset, DepecherError, DeviceError, Engine, PenaltyOpt, declared_symbols, load_library  # noqa: F401

__all__ = ["Engine", "Dataset", "PenaltyOpt", "DepecherError", "DeviceError", "ARI_EXACT", "ARI_MONTE_CARLO", "declared_symbols", "load_library"]
