#!/usr/bin/env python3
"""Benchmark of the per-frame particle physics step (BASELINE.json's metric: particle-steps/s).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one frame (physics -> count -> scan -> pack) over the whole world.
N = 1 : BASELINE.json configs[2], 16 M particles uniform on 5464 x 4096 (the HBM-roofline config).
N > 1 : BASELINE.json configs[4], the 256 M-particle wide world cut into N strips of cell columns,
        one process per GPU (torchrun), edge columns exchanged over NCCL every frame ("strong").
Prints ONE JSON line (rank 0).  `value` = inputs resident in HBM, CUDA-event timed on the worker's
stream.  `e2e` = the same frames driven through the plugin-facing calls with HOST buffers: every
frame uploads the packed frame from pinned memory (write_slice x3 + settings), steps once and reads
the three CPU-visible buffers back at full capacity (plugin/build.rs:88-158 of the reference).
The oracle is only ever run as the CPU baseline here, never on the measured CUDA path.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 8] or \
               [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


def pinned_array(lib, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.wrach_cuda_alloc_host(max(n, 1))
    if not ptr:
        raise MemoryError("cudaMallocHost failed")
    buf = (ctypes.c_char * n).from_address(ptr)
    return np.frombuffer(buf, dtype).reshape(shape), ptr


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port, all host threads): cpu_baseline leg and --impl reference

def cpu_reference(workload, steps, warmup, budget_s, neighbours=False):
    """Times the oracle (OpenMP) on `workload`, shrinking the world (same density) if the requested
    steps would not fit in budget_s.  Returns (value p-s/s, ms per step, cpu_baseline dict)."""
    from oracle import oracle as O
    from wrach_b200 import scene
    wl = scene.WORKLOADS[workload]
    n, dims = wl["n"], wl["dims"]
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1: do not rely on it)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    shrink = 1
    while n // (shrink * shrink) > (1 << 26):  # bounded sample: never build more than 64 M particles on the host
        shrink *= 2
    while True:
        ns, ds = n // (shrink * shrink), (dims[0] // shrink, dims[1] // shrink)
        ow = O.OracleWorld(ds, 3, capacity=int(ns * 1.5) + 64 if wl["pile"] else None, neighbours=neighbours)
        ow.add_particles(O.generate_scene(ns, ds[0], ds[1], seed=scene.SEED, pile=wl["pile"]))
        t0 = time.perf_counter()
        ow.step(1, threads=threads)
        one = time.perf_counter() - t0
        if one * (steps + warmup) <= budget_s or ns <= (1 << 18):
            break
        shrink *= 2
    for _ in range(max(warmup - 1, 0)):
        ow.step(1, threads=threads)
    t0 = time.perf_counter()
    ow.step(steps, threads=threads)
    dt = time.perf_counter() - t0
    value = ow.n * steps / dt
    sample = "%s scene at %d particles on %dx%d (%s), %d frames, oracle OpenMP port, arith=spv" % (
        workload, ow.n, ds[0], ds[1], "full size" if shrink == 1 else "same density, 1/%d area" % (shrink * shrink),
        steps)
    return value, dt / steps * 1e3, {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}


def run_reference(args, rank):
    if rank != 0:
        return
    workload = args.workload or ("16m" if args.gpus == 1 else "256m")
    value, ms, base = cpu_reference(workload, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "description": WORKLOAD_TEXT[workload], "note": "reference physics on host CPU cores (C restatement; "
                       "the Rust/wgpu reference cannot be built in this image)"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# the CUDA path

WORKLOAD_TEXT = {
    "1m-scene": "BASELINE configs[0]: youre-a-pixel scene at 1 M particles (1480x1052)",
    "1m": "BASELINE configs[1]: 1 M particles uniform (1366x1024; L2-resident, not a roofline point)",
    "16m": "BASELINE configs[2]: 16 M particles uniform on 5464x4096 (the HBM-roofline config)",
    "64m-pile": "BASELINE configs[3]: 64 M particles, collapsing pile y = H*u^4 on 10928x8192",
    "256m": "BASELINE configs[4]: 256 M particles uniform on the wide 65532x5462 world",
}


def build_world(W, scene, workload, device, columns=None, strip=None, particles=None):
    """Scene -> WrachState (packed on the host) -> worker (uploaded through the plugin call)."""
    wl = scene.WORKLOADS[workload]
    n, dims = wl["n"], wl["dims"]
    state = W.WrachState(W.WrachConfig(dims, cell_size=3), columns=columns)
    if particles is None:
        particles = scene.generate_fast(n, dims[0], dims[1], pile=wl["pile"])
    state.add_particles(particles)
    del particles
    (gx, gy), total_cells, capacity = state.grid()
    s0 = state.shader_settings
    n_frame = s0.particles_in_frame_count
    capacity = max(capacity, n_frame + (n_frame // 4 + 1024 if strip else 0))  # pile scenes exceed cells*cs^2*1.1
    create = s0.copy()
    create.particles_in_frame_count = 0
    worker = W.PhysicsComputeWorker(create, 0 if strip else total_cells, capacity, device=device, strip=strip)
    W.maybe_upload_to_gpu(worker, state)
    worker.sync()
    return state, worker, dict(n=n_frame, dims=dims, grid=(gx, gy), cells=total_cells - 2, total_cells=total_cells,
                               capacity=capacity, pile=wl["pile"])


def verify_frame(scene, worker, info, columns=None):
    """Read the packed frame back and run the size-independent checks (wrach_host_check_packed) plus
    the order-sensitive checksum.  Not timed."""
    from wrach_b200 import Buffers
    ind = worker.read_vec(Buffers.INDICES_MAIN)
    n = int(ind[-1])
    pos = np.empty((max(n, 1), 2), np.float32)
    vel = np.empty((max(n, 1), 2), np.float32)
    if n:
        worker.read_slice(Buffers.POSITIONS_IN, pos[:n])
        worker.read_slice(Buffers.VELOCITIES_IN, vel[:n])
    gx = info["global_gx"] if "global_gx" in info else info["grid"][0]
    cols = columns or (0, gx)
    scene.check_packed_invariants(ind, pos, vel, cols, gx, info["dims"])
    return n, scene.state_checksum(ind, pos, vel, cols, gx)


def measure_resident(W, scene, workload, steps, warmup, device, peak, keep=False, profile_steps=30, neighbours=False):
    """One BASELINE config on one GPU with the inputs resident: CUDA-event time of `steps` frames on
    the worker's stream after `warmup` frames, the per-kernel split, and the self-checks of the
    frame it ends on."""
    state, worker, info = build_world(W, scene, workload, device)
    if neighbours:  # opt-in extension, not the reference's physics: never the headline line
        worker.set_neighbour_mode(True)
    warm = max(warmup, 3)
    worker.step_timed(warm)
    launches0 = worker.stats()["kernel_launches"]
    sampler = ClockSampler(device)
    sampler.start()
    time.sleep(0.3)
    t0 = time.time()
    ms = worker.step_timed(steps)
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = worker.stats()["kernel_launches"] - launches0
    n_after, checksum = verify_frame(scene, worker, info)
    if n_after != info["n"]:
        raise SystemExit("%s: %d particles after %d frames, %d uploaded" % (workload, n_after, warm + steps, info["n"]))
    prof = max(1, min(steps, profile_steps))
    phys_ms, rebin_ms = worker.step_profiled(prof)
    phys_ms, rebin_ms = phys_ms / prof, rebin_ms / prof
    ab = scene.algorithmic_bytes(info["n"], info["cells"])
    ms_per_step = ms / steps
    st = worker.stats()
    res = {"workload": workload, "description": WORKLOAD_TEXT[workload], "particles": info["n"], "cells": info["cells"],
           "capacity": info["capacity"], "steps": steps, "warmup": warm, "ms_per_step": ms_per_step,
           "value": info["n"] * steps / (ms * 1e-3), "unit": UNIT,
           "step_bytes": ab["step"], "step_gbs": ab["step"] / ms_per_step / 1e6,
           "step_frac": ab["step"] / ms_per_step / 1e6 / peak,
           "k_phys_ms": phys_ms, "k_rebin_ms": rebin_ms, "gpu_launches": int(launches), "clocks": clocks,
           "slow_path_frames": st["slow_path_steps"],
           "path": ("k_tile_frame (one fused launch per frame over 22x14-cell tiles)" if st["tile_frames"] and not st["tile_fallbacks"]
                    else "k_phys + k_run_scan + k_rebin" + (" (after %d tile fall-backs)" % st["tile_fallbacks"] if st["tile_fallbacks"] else "")),
           "tile_frames": st["tile_frames"], "tile_fallbacks": st["tile_fallbacks"],
           "verified": "N conserved, indices monotone, every particle in the slot range of its cell, inside the world, |v|<=1",
           "state_checksum": "%016x" % checksum, "frames_at_checksum": warm + steps}
    if keep:
        return res, state, worker, info, (phys_ms, rebin_ms, ab)
    worker.close()
    state.close()
    return res


def run_single(args):
    import wrach_b200 as W
    from wrach_b200 import Buffers, _ffi, scene
    workload = args.workload or "16m"
    peak, peak_src = load_peaks()
    lib = _ffi.lib()

    # ---- value: resident inputs, CUDA events on the worker's stream
    res, state, worker, info, (phys_ms, rebin_ms, ab) = measure_resident(
        W, scene, workload, args.steps, args.warmup, args.device, peak, keep=True, profile_steps=50, neighbours=args.neighbours)
    n_frame, total_cells, capacity, cells = info["n"], info["total_cells"], info["capacity"], info["cells"]
    ms_per_step, value = res["ms_per_step"], res["value"]
    if rebin_ms == 0.0:
        # fused tile frames: ONE launch does the whole step, so its algorithmic bytes are the step's
        # (64 N + 16 C); it really moves about half of that (see `traffic`)
        kernels = {"k_tile_frame": {"ms": phys_ms, "bytes": ab["step"]}}
    else:
        kernels = {"k_phys": {"ms": phys_ms, "bytes": ab["phys"]}, "k_rebin": {"ms": rebin_ms, "bytes": ab["rebin"]}}
    for k in kernels.values():
        k["gbs"] = k["bytes"] / k["ms"] / 1e6
        k["frac"] = k["gbs"] / peak
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    dom_ms, dom_bytes = kernels[dom]["ms"], kernels[dom]["bytes"]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                "kernels": kernels,
                "step": {"bytes": ab["step"], "gbs": res["step_gbs"], "frac": res["step_frac"]}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)
    if os.path.exists(traffic_file):  # dram bytes per launch from the committed ncu --set full capture
        try:
            t = json.load(open(traffic_file))
            roofline["traffic"] = t.get(dom)
            roofline["traffic_source"] = "profiles/traffic_%s.json (%s), not measured in this run" % (
                workload, t.get("source", "ncu --set full capture of the same command"))
        except Exception:
            pass

    # ---- e2e: host buffers, copies inside the timed region, through the plugin-facing calls
    ind_h, p1 = pinned_array(lib, (total_cells,), np.uint32)
    pos_h, p2 = pinned_array(lib, (capacity, 2), np.float32)
    vel_h, p3 = pinned_array(lib, (capacity, 2), np.float32)
    e2e_steps = max(3, min(args.steps, 10))
    h2d = n_frame * 16 + total_cells * 4 + 32
    d2h = capacity * 16 + total_cells * 4
    settings = worker.settings.copy()

    def read_back(n_slots):
        # tick (plugin/build.rs:135-158): three read_vec copies queued, ONE synchronisation
        worker.read_slice_async(Buffers.INDICES_MAIN, ind_h)
        worker.read_slice_async(Buffers.POSITIONS_IN, pos_h[:n_slots])
        worker.read_slice_async(Buffers.VELOCITIES_IN, vel_h[:n_slots])
        worker.sync()

    def frame():
        # maybe_upload_to_gpu with a pending GPUUpload::PackedData + Settings (plugin/build.rs:88-126)
        worker.write_slice(Buffers.INDICES_MAIN, ind_h)
        worker.write_slice(Buffers.POSITIONS_IN, pos_h[:n_frame])
        worker.write_slice(Buffers.VELOCITIES_IN, vel_h[:n_frame])
        worker.write(Buffers.WORLD_SETTINGS_UNIFORM, settings)
        worker.step(1)
        read_back(capacity)

    # the PCIe wall on this box: the position buffer alone, pinned, each way (best of 3)
    read_back(capacity)
    pcie = {}
    for name, fn in (("d2h_gbs", lambda: worker.read_slice(Buffers.POSITIONS_IN, pos_h)),
                     ("h2d_gbs", lambda: (worker.write_slice(Buffers.POSITIONS_IN, pos_h), worker.sync()))):
        best = 1e9
        for _ in range(3):
            tq = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - tq)
        pcie[name] = capacity * 8 / best / 1e9
    read_back(capacity)
    for _ in range(2):
        frame()
    worker.sync()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        frame()
    worker.sync()
    e2e_dt = time.perf_counter() - te
    wall = h2d / (pcie["h2d_gbs"] * 1e9) + d2h / (pcie["d2h_gbs"] * 1e9) + ms_per_step * 1e-3
    e2e = {"value": n_frame * e2e_steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_dt / e2e_steps * 1e3,
           "pcie_gbs": (h2d + d2h) / (e2e_dt / e2e_steps) / 1e9, "pcie_pinned_copy": pcie,
           "frac_of_copy_wall": wall / (e2e_dt / e2e_steps),
           "path": "write_slice x3 + write(settings) + step(1) + read_async x3 (capacity-sized) + one sync, pinned host memory"}
    # the reference's steady-state frame: nothing new to upload (state stays resident), one step, the
    # plugin's own tick (C++ host mirror: three capacity-sized read-backs into WrachState.packed_data)
    W.tick(worker, state)
    worker.sync()
    tt = time.perf_counter()
    for _ in range(e2e_steps):
        worker.write(Buffers.WORLD_SETTINGS_UNIFORM, settings)
        worker.step(1)
        W.tick(worker, state)
    tick_dt = time.perf_counter() - tt
    e2e["tick_only"] = {"value": n_frame * e2e_steps / tick_dt, "h2d_bytes_per_step": 32, "d2h_bytes_per_step": d2h,
                        "ms_per_step": tick_dt / e2e_steps * 1e3, "d2h_gbs": d2h / (tick_dt / e2e_steps) / 1e9,
                        "frac_of_pinned_d2h": (d2h / (tick_dt / e2e_steps) / 1e9) / pcie["d2h_gbs"],
                        "path": "write(settings) + step(1) + wrach_plugin_tick_wait: state resident, as WrachAPI::tick does"}
    # ... and with the read-back cut to the N live particles (wrach_plugin_tick_active)
    worker.sync()
    ta = time.perf_counter()
    for _ in range(e2e_steps):
        worker.write(Buffers.WORLD_SETTINGS_UNIFORM, settings)
        worker.step(1)
        W.tick_active(worker, state)
    act_dt = time.perf_counter() - ta
    e2e["tick_active"] = {"value": n_frame * e2e_steps / act_dt, "h2d_bytes_per_step": 32,
                          "d2h_bytes_per_step": n_frame * 16 + total_cells * 4,
                          "ms_per_step": act_dt / e2e_steps * 1e3,
                          "path": "as tick_only, but the read-back takes the N live slots (tick_active), not the capacity"}
    worker.close()
    state.close()
    for p in (p1, p2, p3):
        lib.wrach_cuda_free_host(p)

    # ---- the reference's CPU path on this box's cores, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        _, _, cpu = cpu_reference(workload, steps=5, warmup=1, budget_s=25.0, neighbours=args.neighbours)

    # ---- the other BASELINE configs, one after the other on the same GPU (resident inputs only)
    extra, scale_base = {}, None
    if not args.no_extra and workload == "16m":
        plan = (("1m-scene", 1000), ("1m", args.steps), ("64m-pile", min(args.steps, 60)), ("256m", args.steps))
        for wl_name, k in plan:
            try:
                extra[wl_name] = measure_resident(W, scene, wl_name, k, args.warmup, args.device, peak)
            except Exception as e:  # never lose the headline to an extra
                extra[wl_name] = {"error": repr(e)}
        if "value" in extra["1m-scene"] and not args.no_cpu_baseline:
            _, _, c0 = cpu_reference("1m-scene", steps=1000, warmup=1, budget_s=45.0)
            extra["1m-scene"]["cpu_baseline"] = c0
            extra["1m-scene"]["readme_fps"] = "reference README: 1,000,000 particles at ~39 fps (unstated hardware, windowed)"
            extra["1m-scene"]["fps_resident"] = 1e3 / extra["1m-scene"]["ms_per_step"]
        if "value" in extra["256m"]:
            b = extra["256m"]
            scale_base = {"workload": "256m", "n_gpus": 1, "ms_per_step": b["ms_per_step"], "value": b["value"],
                          "state_checksum": b["state_checksum"], "frames_at_checksum": b["frames_at_checksum"],
                          "note": "the world bench.py --gpus N>1 cuts into strips, on this one GPU, same run"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "description": "%s: %d particles%s on %dx%d, cell 3, grid %dx%d (%d cells), capacity %d" % (
                WORKLOAD_TEXT[workload], n_frame, " pile y=H*u^4" if info["pile"] else "", info["dims"][0], info["dims"][1],
                info["grid"][0], info["grid"][1], cells, capacity),
                "seed": hex(scene.SEED), "arith": "spv", "l2": "working set %.2f GB > 126 MB L2, no flush needed" % (
                    (n_frame * 33 * 2 + total_cells * 8) / 1e9),
                "slow_path_frames": res["slow_path_frames"], "path": res["path"], "tile_frames": res["tile_frames"],
                "tile_fallbacks": res["tile_fallbacks"], "verified": res["verified"],
                "state_checksum": res["state_checksum"], "frames_at_checksum": res["frames_at_checksum"],
                **({"mode": "3x3 neighbour pass before every frame (extension, not in the reference)"} if args.neighbours else {})},
            "clocks": res["clocks"], "e2e": e2e, "gpu_launches": res["gpu_launches"], "roofline": roofline, "cpu_baseline": cpu,
            "extra_configs": extra, "scale_base": scale_base}
    emit(line)


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner when NCCL_DEBUG is set on the box), so everything that goes to file descriptor 1 during
    the run is sent to stderr, and emit() writes the JSON line to the descriptor saved here (kept in
    the environment: bench_strips imports this file as a second module object)."""
    sys.stdout.flush()
    os.environ["WRACH_BENCH_STDOUT_FD"] = str(os.dup(1))
    os.dup2(2, 1)


def emit(line):
    os.write(int(os.environ.get("WRACH_BENCH_STDOUT_FD", "1")), (json.dumps(line) + "\n").encode())


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="1m-scene | 1m | 16m | 64m-pile | 256m")
    ap.add_argument("--device", type=int, default=int(os.environ.get("LOCAL_RANK", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="N=1: skip the other BASELINE configs (extra_configs, scale_base)")
    ap.add_argument("--no-scale-base", action="store_true", help="N>1: skip the same-world 1-GPU point measured by rank 0")
    ap.add_argument("--neighbours", action="store_true",
                    help="N=1 only: time the opt-in 3x3 neighbour mode (an extension; the default line is the reference's physics)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.gpus > 1:
        from bench_strips import run_strips  # one process per GPU, launched by torchrun
        run_strips(args)
        return
    run_single(args)


if __name__ == "__main__":
    main()
