#!/usr/bin/env python3
"""GPU box: device-to-host copy rate into cudaMallocHost, registered and pageable memory, and what the
plugin's tick (three capacity-sized read-backs into the state's own arrays) achieves next to them."""
import sys, time, ctypes, numpy as np
sys.path.insert(0, "/root/repo")
import wrach_b200 as W
from wrach_b200 import Buffers, scene, _ffi
wl = scene.WORKLOADS["16m"]
state = W.WrachState(W.WrachConfig(wl["dims"], cell_size=3))
state.add_particles(scene.generate_fast(wl["n"], wl["dims"][0], wl["dims"][1]))
(gx, gy), total_cells, capacity = state.grid()
s = state.shader_settings
create = s.copy(); create.particles_in_frame_count = 0
w = W.PhysicsComputeWorker(create, total_cells, capacity)
W.maybe_upload_to_gpu(w, state)
w.step(2); w.sync()
L = _ffi.lib()
nb = capacity * 8
def bench(ptr, label, reps=10):
    L.wrach_cuda_read(w._h, _ffi.POSITIONS_IN, ctypes.c_void_p(ptr), ctypes.c_size_t(nb))
    t = time.perf_counter()
    for _ in range(reps): L.wrach_cuda_read(w._h, _ffi.POSITIONS_IN, ctypes.c_void_p(ptr), ctypes.c_size_t(nb))
    dt = (time.perf_counter() - t) / reps
    print("%-28s %.3f ms  %.1f GB/s" % (label, dt * 1e3, nb / dt / 1e9), flush=True)
p1 = L.wrach_cuda_alloc_host(nb)
bench(p1, "cudaMallocHost")
a = np.empty(nb, np.uint8); a[:] = 0
L.wrach_cuda_host_register(a.ctypes.data, nb)
bench(a.ctypes.data, "numpy + cudaHostRegister")
b = np.empty(nb, np.uint8); b[:] = 0
bench(b.ctypes.data, "pageable")
# the plugin's tick: step(1) + three read-backs into the state's own vectors
from wrach_b200 import api
for _ in range(3):
    w.step(1); api.tick(w, state, wait=True)
t = time.perf_counter()
for _ in range(8):
    w.step(1); api.tick(w, state, wait=True)
dt = (time.perf_counter() - t) / 8
print("step(1) + tick: %.3f ms" % (dt * 1e3))
t = time.perf_counter()
for _ in range(8):
    api.tick(w, state, wait=True)
dt = (time.perf_counter() - t) / 8
print("tick alone (packed valid): %.3f ms  %.1f GB/s" % (dt * 1e3, (2 * nb + (total_cells) * 4) / dt / 1e9))
