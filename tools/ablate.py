#!/usr/bin/env python3
"""Debug tool: time k_phys alone (wrach_cuda_debug_phys_only, -DWRACH_DEBUG_PHYS_ONLY builds) for every
library in wrach_b200/lib/ablate/.  Variants built with -DWRACH_ABLATE=<bits> drop parts of the kernel
(results are wrong; only the time is of interest).  Run on the GPU box:
    python tools/ablate.py [workload]        (spawns one process per variant)"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(lib_path, workload):
    os.environ["WRACH_CUDA_LIB"] = lib_path
    import wrach_b200 as W
    from wrach_b200 import _ffi, scene

    wl = scene.WORKLOADS[workload]
    state = W.WrachState(W.WrachConfig(wl["dims"], cell_size=3))
    (gx, gy), total_cells, capacity = state.grid()
    state.add_particles(scene.generate_fast(wl["n"], *wl["dims"], pile=wl["pile"]))
    s0 = state.shader_settings.copy()
    s0.particles_in_frame_count = 0
    worker = W.PhysicsComputeWorker(s0, total_cells, max(capacity, wl["n"]))
    W.maybe_upload_to_gpu(worker, state)
    worker.sync()
    L = _ffi.lib()
    L.wrach_cuda_debug_phys_only.restype = ctypes.c_int
    L.wrach_cuda_debug_phys_only.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float)]
    ms = ctypes.c_float()
    res = []
    for _ in range(3):
        assert L.wrach_cuda_debug_phys_only(worker._h, 50, ctypes.byref(ms)) == 0
        res.append(ms.value)
    print("%-28s k_phys %.4f ms  (%s)" % (os.path.basename(lib_path), min(res), " ".join("%.4f" % r for r in res)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2], sys.argv[3])
    else:
        wl = sys.argv[1] if len(sys.argv) > 1 else "16m"
        for lib in sorted(glob.glob(os.path.join(ROOT, "wrach_b200", "lib", "ablate", "*.so"))):
            subprocess.run([sys.executable, __file__, "--one", lib, wl], check=False)
