"""wrach_b200 — B200-native (sm_100a) CUDA implementation of Wrach's per-frame particle physics
step behind the reference's compute-worker boundary.  The compiled library
(wrach_b200/lib/libwrach_cuda.so) is the product; importing the worker without it raises."""
from ._ffi import (ARITH_SPV, ARITH_UNFUSED, LIB_PATH, WorldSettings, WrachCudaError)  # noqa: F401
from .worker import Buffers, PhysicsComputeWorker  # noqa: F401
from .api import (WrachAPI, WrachConfig, WrachState, active_grid, get_active_cells, get_cell_coord,  # noqa: F401
                  max_particles_per_frame, maybe_upload_to_gpu, tick, tick_active)
