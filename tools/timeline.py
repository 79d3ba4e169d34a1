#!/usr/bin/env python3
"""Debug tool: per-phase timeline of one k_rebin launch from the -DWRACH_TIMELINE build
(wrach_b200/lib/libwrach_cuda_timeline.so).  Run on the GPU box:  python tools/timeline.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wrach_b200 import _ffi  # noqa: E402

_ffi.LIB_PATH = os.path.join(ROOT, "wrach_b200", "lib", "libwrach_cuda_timeline.so")
import wrach_b200 as W  # noqa: E402
from wrach_b200 import scene  # noqa: E402

wl = scene.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "16m"]
state = W.WrachState(W.WrachConfig(wl["dims"], cell_size=3))
(gx, gy), total_cells, capacity = state.grid()
state.add_particles(scene.generate_fast(wl["n"], *wl["dims"], pile=wl["pile"]))
s0 = state.shader_settings.copy()
s0.particles_in_frame_count = 0
worker = W.PhysicsComputeWorker(s0, total_cells, max(capacity, wl["n"]))
W.maybe_upload_to_gpu(worker, state)
worker.step(10)
worker.sync()
L = _ffi.lib()
nb = (gx * gy + 255) // 256
npb = (gx * gy + 255) // 256  # k_phys blocks are stamped behind the k_rebin tiles... at [gridDim_phys + block]
buf = np.zeros((nb + 2 * npb + 16, 16), np.uint64)
L.wrach_cuda_debug_timeline.restype = ctypes.c_int
L.wrach_cuda_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
assert L.wrach_cuda_debug_timeline(worker._h, buf.ctypes.data, buf.shape[0]) == 0
tp = buf[npb:2 * npb].astype(np.float64)  # k_phys stamps live at [gridDim.x + blockIdx.x]
t = buf[:nb].astype(np.float64)
pn = ["start->starts loaded", "sort+cell ids", "wait for TMA", "pairs (warp 0)", "pairs barrier", "finish"]
print("k_phys blocks", npb, " span %.1f us" % ((tp[:, 6].max() - tp[:, 0].min()) / 1e3))
lifep = (tp[:, 6] - tp[:, 0]) / 1e3
print("block lifetime us: mean %.2f p50 %.2f p90 %.2f max %.2f" % (lifep.mean(), np.median(lifep), np.percentile(lifep, 90), lifep.max()))
for i, n in enumerate(pn):
    d = (tp[:, i + 1] - tp[:, i]) / 1e3
    print("  %-22s mean %6.2f us  p50 %6.2f  p90 %6.2f  max %7.2f" % (n, d.mean(), np.median(d), np.percentile(d, 90), d.max()))
t0 = t[:, 1:9][t[:, 1:9] > 0].min()
names = ["-", "starts+cls+list sizes", "entries+tma wait", "-", "rank+count+scan", "tables", "copy row", "copy vertical"]
print("tiles", nb, " kernel span %.1f us" % ((t[:, 8].max() - t0) / 1e3))
life = (t[:, 8] - t[:, 1]) / 1e3
print("tile lifetime us: mean %.2f  p50 %.2f  p90 %.2f  max %.2f" % (life.mean(), np.median(life), np.percentile(life, 90), life.max()))
for i, n in enumerate(names):
    a, b = (3, 5) if i == 4 else (i, i + 1)
    d = (t[:, b] - t[:, a]) / 1e3
    if n == "-":
        continue
    print("  %-18s mean %6.2f us  p50 %6.2f  p90 %6.2f  max %7.2f" % (n, d.mean(), np.median(d), np.percentile(d, 90), d.max()))
for a, b, n in ((3, 9, "  rank_vertical (thread 0)"), (9, 10, "  barrier"), (10, 11, "  size tables"), (11, 5, "  block scan")):
    d = (t[:, b] - t[:, a]) / 1e3
    print("  %-26s mean %6.2f us  p50 %6.2f  p90 %6.2f" % (n, d.mean(), np.median(d), np.percentile(d, 90)))
start = (t[:, 1] - t0) / 1e3
print("tile start time us (by tile id) deciles:", np.round(np.percentile(start, [0, 10, 25, 50, 75, 90, 100]), 1))
order_violation = np.sum(np.diff(t[:, 1]) < 0)
print("tiles whose start precedes their predecessor's:", int(order_violation))
# look-back: stamp 5 = the run's total is known, stamp 6 = its first slot is known (all predecessors summed)
wait = (t[:, 6] - t[:, 5]) / 1e3
ready = np.maximum.accumulate(t[:, 5])  # earliest moment every predecessor's total exists
lag = (t[:, 6] - ready) / 1e3
print("look-back: wait mean %.2f p50 %.2f p90 %.2f max %.2f us;  lag behind the slowest predecessor mean %.2f p90 %.2f us" % (
    wait.mean(), np.median(wait), np.percentile(wait, 90), wait.max(), lag.mean(), np.percentile(lag, 90)))
own = (t[:, 5] >= ready)
print("tiles that were themselves the slowest so far: %d of %d" % (int(own.sum()), nb))
sp = (t[:, 5] - t0) / 1e3
print("scan-point time by tile id, deciles:", np.round(np.percentile(sp, [0, 10, 25, 50, 75, 90, 100]), 1))
# the slowest blocks (skewed scenes: a dense run is one block's job)
for name, arr, a, b in (("k_phys", tp, 0, 6), ("k_rebin", t, 1, 8)):
    lf = (arr[:, b] - arr[:, a]) / 1e3
    top = np.argsort(lf)[-5:][::-1]
    print(name, "slowest blocks:", ", ".join("%d: %.0f us" % (i, lf[i]) for i in top),
          " | sum of lifetimes %.0f us over %d SMs x resident blocks" % (lf.sum(), 148))
