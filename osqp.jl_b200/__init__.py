"""B200-native OSQP ADMM engine -- host-side mirror of the reference's API.

Only what the hot path needs lives here:
  csrc/        CUDA kernels (sm_100a) + the C ABI of include/osqp.h  -> lib/libosqp.so
  interface.py mirror of src/interface.jl over ctypes (Model/setup/solve/update/warm_start)
  types.py     mirror of src/types.jl (struct layouts)
  constants.py mirror of src/constants.jl
  batch.py     batched extension (many small independent QPs, sharded across GPUs)
"""
from .constants import *  # noqa: F401,F403
from .interface import ABI_SYMBOLS, DEFAULT_LIB, Model, ccsc_to_scipy, load_library, ManagedCcsc  # noqa: F401
from .types import Ccsc, CInfo, Data, Info, Results, Settings, Solution, Workspace  # noqa: F401
from . import types  # noqa: F401
from .batch import BatchModel, BatchResults, shard_range  # noqa: F401
