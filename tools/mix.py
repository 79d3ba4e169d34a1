#!/usr/bin/env python3
"""Debug tool (build with -DWRACH_DEBUG_MIX into wrach_b200/lib/libwrach_cuda_mix.so): what does one
launch of alternating re-bin and physics blocks (of two independent frames) cost against the two
kernels back to back?  Run on the GPU box:  python tools/mix.py [workload]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WRACH_CUDA_LIB"] = os.path.join(ROOT, "wrach_b200", "lib", os.environ.get("MIXLIB", "libwrach_cuda_mix.so"))
import wrach_b200 as W  # noqa: E402
from wrach_b200 import _ffi, scene  # noqa: E402

wl = scene.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "16m"]
workers = []
for seed in (scene.SEED, scene.SEED + 1):
    state = W.WrachState(W.WrachConfig(wl["dims"], cell_size=3))
    (gx, gy), total_cells, capacity = state.grid()
    state.add_particles(scene.generate_fast(wl["n"], *wl["dims"], seed=seed, pile=wl["pile"]))
    s0 = state.shader_settings.copy()
    s0.particles_in_frame_count = 0
    w = W.PhysicsComputeWorker(s0, total_cells, max(capacity, wl["n"]))
    W.maybe_upload_to_gpu(w, state)
    w.step(5)
    w.sync()
    workers.append(w)
L = _ffi.lib()
L.wrach_cuda_debug_mix.restype = ctypes.c_int
L.wrach_cuda_debug_mix.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float),
                                   ctypes.POINTER(ctypes.c_float)]
a, b = ctypes.c_float(), ctypes.c_float()
for _ in range(3):
    assert L.wrach_cuda_debug_mix(workers[0]._h, workers[1]._h, 30, ctypes.byref(a), ctypes.byref(b)) == 0
    print("k_rebin ; k_phys back to back %.4f ms   one mixed launch %.4f ms   (%.1f %%)" % (a.value, b.value, 100 * b.value / a.value))
