"""N > 1 leg of bench.py: BASELINE.json configs[4], the wide 256 M-particle world cut into N strips
of cell columns, one process per GPU (torchrun), edge-column particles exchanged by ncclSend/ncclRecv
on each worker's own stream every frame.  Total work is fixed as N grows ("strong").

torch is only plumbing here: process group, barrier, max-over-ranks of the per-rank CUDA-event
times and the broadcast of the ncclUniqueId.  Every kernel on the timed path is this repo's.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def strip_scene(scene, W, wl, rank, world, gx, cell=3):
    """Particles of this rank's strip: uniform density, x inside the strip's columns.  Counts are
    proportional to strip width and sum to wl['n'] exactly; ids are disjoint across ranks."""
    n, (width, height) = wl["n"], wl["dims"]
    edges = [min(float(width), float(cell * W.PhysicsComputeWorker.strip_columns(gx, r, world)[0])) for r in range(world)]
    edges.append(float(width))
    cum = [int(round(n * e / width)) for e in edges]
    cum[-1] = n
    x0, x1 = edges[rank], edges[rank + 1]
    return scene.generate_fast(cum[rank + 1] - cum[rank], x1 - x0, height, first_id=cum[rank], pile=wl["pile"], x0=x0)


def run_strips(args):
    import torch
    import torch.distributed as dist

    import bench
    import wrach_b200 as W
    from wrach_b200 import Buffers, _ffi, scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    workload = args.workload or "256m"
    wl = scene.WORKLOADS[workload]
    dims = wl["dims"]
    peak, peak_src = bench.load_peaks()
    lib = _ffi.lib()

    config = W.WrachConfig(dims, cell_size=3)
    _, (gx, gy) = W.active_grid((0.0, 0.0, dims[0], dims[1]), 3)
    cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
    state = W.WrachState(config, columns=cols)
    particles = strip_scene(scene, W, wl, rank, world, gx)
    state.add_particles(particles)
    del particles
    gsettings = state.shader_settings.copy()
    n_local = gsettings.particles_in_frame_count
    _, total_cells, capacity = state.grid()
    capacity = max(capacity, int(n_local * 1.25) + 1024)
    cells_local = total_cells - 2

    # one ncclUniqueId for the strip communicator, made by rank 0
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(W.PhysicsComputeWorker.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    create = gsettings.copy()
    create.particles_in_frame_count = 0
    worker = W.PhysicsComputeWorker(create, 0, capacity, device=local_rank, strip=(rank, world, bytes(uid.cpu().numpy())))
    W.maybe_upload_to_gpu(worker, state)
    worker.sync()

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n_total = int(all_sum(n_local))
    cells_total = int(all_sum(cells_local))

    # ---- value: resident inputs, per-rank CUDA events on the worker's stream, max over ranks
    worker.step_timed(max(args.warmup, 3))
    launches0 = worker.stats()["kernel_launches"]
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    ms_local = worker.step_timed(args.steps)
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.time()
    ms = all_max(ms_local)
    clocks = sampler.stop(t0, t1) if sampler else None
    st = worker.stats()
    launches = int(all_sum(st["kernel_launches"] - launches0))
    halo = st["halo_bytes_sent"]
    value = n_total * args.steps / (ms * 1e-3)

    prof_steps = min(args.steps, 30)
    phys_ms, rebin_ms = worker.step_profiled(prof_steps)
    phys_ms, rebin_ms = all_max(phys_ms / prof_steps), all_max(rebin_ms / prof_steps)
    ab = scene.algorithmic_bytes(n_local, cells_local)
    dom = "k_phys" if phys_ms >= rebin_ms else "k_rebin"
    dom_ms, dom_bytes = (phys_ms, ab["phys"]) if dom == "k_phys" else (rebin_ms, ab["rebin"])

    # ---- e2e: every rank uploads its packed strip from pinned memory, steps once, reads the three
    # CPU-visible buffers back at full capacity (plugin/build.rs:88-158), max over ranks
    ind_h, p1 = bench.pinned_array(lib, (total_cells,), np.uint32)
    pos_h, p2 = bench.pinned_array(lib, (capacity, 2), np.float32)
    vel_h, p3 = bench.pinned_array(lib, (capacity, 2), np.float32)
    e2e_steps = max(3, min(args.steps, 6))
    settings = gsettings.copy()

    def frame():
        worker.write_slice(Buffers.INDICES_MAIN, ind_h)
        worker.write_slice(Buffers.POSITIONS_IN, pos_h[:n_now[0]])
        worker.write_slice(Buffers.VELOCITIES_IN, vel_h[:n_now[0]])
        settings.particles_in_frame_count = n_now[0]
        worker.write(Buffers.WORLD_SETTINGS_UNIFORM, settings)
        worker.step(1)
        worker.read_vec(Buffers.INDICES_MAIN, out=ind_h)
        worker.read_vec(Buffers.POSITIONS_IN, out=pos_h)
        worker.read_vec(Buffers.VELOCITIES_IN, out=vel_h)
        n_now[0] = int(ind_h[-1])  # particles migrate: the strip's count changes every frame

    worker.read_vec(Buffers.INDICES_MAIN, out=ind_h)
    worker.read_vec(Buffers.POSITIONS_IN, out=pos_h)
    worker.read_vec(Buffers.VELOCITIES_IN, out=vel_h)
    n_now = [int(ind_h[-1])]
    frame()
    dist.barrier()
    torch.cuda.synchronize()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        frame()
    worker.sync()
    e2e_dt = all_max(time.perf_counter() - te)
    h2d = all_sum(n_local * 16 + total_cells * 4 + 32)
    d2h = all_sum(capacity * 16 + total_cells * 4)
    worker.close()
    for p in (p1, p2, p3):
        lib.wrach_cuda_free_host(p)

    if rank == 0:
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        step_bytes = scene.algorithmic_bytes(n_total, cells_total)["step"]
        line = {"metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s: %d particles uniform on %dx%d, cell 3, grid %dx%d, %d strips of cell columns "
                                       "(one per GPU), edge columns exchanged over NCCL every frame" % (
                                           workload, n_total, dims[0], dims[1], gx, gy, world),
                           "seed": hex(scene.SEED), "arith": "spv",
                           "l2": "per-GPU working set %.2f GB > 126 MB L2, no flush needed" % (n_local * 68 / 1e9),
                           "halo_bytes_per_step_per_rank": halo // max(1, st["steps_completed"])},
                "clocks": clocks,
                "e2e": {"value": n_total * e2e_steps / e2e_dt, "unit": bench.UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                        "path": "per rank: write_slice x3 + write(settings) + step(1) + read_vec x3 (capacity-sized), pinned"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "per": "one GPU (max over ranks)",
                             "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                             "kernels": {"k_phys": {"ms": phys_ms, "bytes": ab["phys"]},
                                         "k_rebin": {"ms": rebin_ms, "bytes": ab["rebin"]}},
                             "step": {"bytes": step_bytes, "gbs_all_gpus": step_bytes / (ms / args.steps) / 1e6,
                                      "frac_of_n_x_peak": step_bytes / (ms / args.steps) / 1e6 / (peak * world)}},
                "cpu_baseline": None}
        bench.emit(line)
    dist.destroy_process_group()
