"""N > 1 leg of bench.py: BASELINE.json configs[4], the wide 256 M-particle world cut into N strips
of cell columns, one process per GPU (torchrun), edge-column particles exchanged over NCCL on each
worker's own streams every frame.  Total work is fixed as N grows ("strong").

Every N runs the SAME 2^28-particle scene (global particle ids, bucketed by cell column: nothing is
generated per strip, nothing is lost on a strip edge).  After the timed frames every rank reads its
strip back and checks it (N conserved over all ranks, indices monotone, every particle in the slot
range of the cell its position keys to, inside the world, |v| <= 1) and the ranks' order-sensitive
checksums are added up.  Rank 0 then runs the whole world alone on its GPU for the same number of
frames (`scale_base`): the line carries `speedup_vs_1gpu_same_world`, and the single-GPU checksum
must equal the strips' sum -- N GPUs and one GPU hold the same bits.

torch is only plumbing here: process group, barrier, max-over-ranks of the per-rank CUDA-event
times and the broadcast of the ncclUniqueId.  Every kernel on the timed path is this repo's.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


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

    _, (gx, gy) = W.active_grid((0.0, 0.0, dims[0], dims[1]), 3)
    cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
    particles = scene.generate_columns(wl["n"], dims[0], dims[1], cols, pile=wl["pile"])

    # one ncclUniqueId for the strip communicator, made by rank 0
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(W.PhysicsComputeWorker.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    state, worker, info = bench.build_world(W, scene, workload, local_rank, columns=cols,
                                            strip=(rank, world, bytes(uid.cpu().numpy())), particles=particles)
    del particles
    info["global_gx"] = gx
    n_local, total_cells, capacity, cells_local = info["n"], info["total_cells"], info["capacity"], info["cells"]
    gsettings = state.shader_settings.copy()

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum_int(x):
        t = torch.tensor([int(x) & 0x7FFFFFFF, int(x) >> 31], dtype=torch.int64, device="cuda")  # exact beyond 2^53
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t[0].item()) + (int(t[1].item()) << 31)

    n_total = all_sum_int(n_local)
    cells_total = all_sum_int(cells_local)
    if n_total != wl["n"]:
        raise SystemExit("the strips hold %d particles, the scene has %d" % (n_total, wl["n"]))

    # ---- value: resident inputs, per-rank CUDA events on the worker's stream, max over ranks
    warm = max(args.warmup, 3)
    worker.step_timed(warm)
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
    launches = all_sum_int(st["kernel_launches"] - launches0)
    halo = st["halo_bytes_sent"]
    value = n_total * args.steps / (ms * 1e-3)

    # ---- the frame the timed region ended on: per-rank self-checks, global N, summed checksum
    n_now, checksum = bench.verify_frame(scene, worker, info, columns=cols)
    n_after = all_sum_int(n_now)
    if n_after != n_total:
        raise SystemExit("%d particles over all strips after %d frames, %d before" % (n_after, warm + args.steps, n_total))
    checksum_all = all_sum_int(checksum) & 0xFFFFFFFFFFFFFFFF

    prof_steps = min(args.steps, 30)
    phys_ms, rebin_ms = worker.step_profiled(prof_steps)
    phys_ms, rebin_ms = all_max(phys_ms / prof_steps), all_max(rebin_ms / prof_steps)
    ab = scene.algorithmic_bytes(n_local, cells_local)
    st2 = worker.stats()
    tiles = st2["tile_frames"] > 0 and rebin_ms == 0.0
    if tiles:  # one fused launch per frame (plus the two edge-column launches): the whole step's bytes
        kernels = {"k_tile_frame": {"ms": phys_ms, "bytes": ab["step"]}}
    else:
        kernels = {"k_phys": {"ms": phys_ms, "bytes": ab["phys"]}, "k_rebin": {"ms": rebin_ms, "bytes": ab["rebin"]}}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    dom_ms, dom_bytes = kernels[dom]["ms"], kernels[dom]["bytes"]

    # ---- e2e: every rank uploads its packed strip from pinned memory, steps once, reads the three
    # CPU-visible buffers back at full capacity (plugin/build.rs:88-158), max over ranks
    ind_h, p1 = bench.pinned_array(lib, (total_cells,), np.uint32)
    pos_h, p2 = bench.pinned_array(lib, (capacity, 2), np.float32)
    vel_h, p3 = bench.pinned_array(lib, (capacity, 2), np.float32)
    e2e_steps = max(3, min(args.steps, 6))
    settings = gsettings.copy()

    def read_back():
        worker.read_slice_async(Buffers.INDICES_MAIN, ind_h)
        worker.read_slice_async(Buffers.POSITIONS_IN, pos_h)
        worker.read_slice_async(Buffers.VELOCITIES_IN, vel_h)
        worker.sync()

    def frame():
        worker.write_slice(Buffers.INDICES_MAIN, ind_h)
        worker.write_slice(Buffers.POSITIONS_IN, pos_h[:n_now_box[0]])
        worker.write_slice(Buffers.VELOCITIES_IN, vel_h[:n_now_box[0]])
        settings.particles_in_frame_count = n_now_box[0]
        worker.write(Buffers.WORLD_SETTINGS_UNIFORM, settings)
        worker.step(1)
        read_back()
        n_now_box[0] = int(ind_h[-1])  # particles migrate: the strip's count changes every frame

    read_back()
    n_now_box = [int(ind_h[-1])]
    frame()
    dist.barrier()
    torch.cuda.synchronize()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        frame()
    worker.sync()
    e2e_dt = all_max(time.perf_counter() - te)
    h2d = all_sum_int(n_local * 16 + total_cells * 4 + 32)
    d2h = all_sum_int(capacity * 16 + total_cells * 4)
    worker.close()
    state.close()
    for p in (p1, p2, p3):
        lib.wrach_cuda_free_host(p)

    # ---- the same world on ONE GPU, same frames, same run (rank 0; the other ranks wait)
    scale_base = None
    if rank == 0 and not args.no_scale_base:
        try:
            b = bench.measure_resident(W, scene, workload, args.steps, args.warmup, local_rank, peak)
            scale_base = {"workload": workload, "n_gpus": 1, "ms_per_step": b["ms_per_step"], "value": b["value"],
                          "step_frac": b["step_frac"], "state_checksum": b["state_checksum"],
                          "frames_at_checksum": b["frames_at_checksum"], "clocks": b["clocks"],
                          "checksum_equals_strips": b["state_checksum"] == "%016x" % checksum_all}
        except Exception as e:
            scale_base = {"error": repr(e)}
    dist.barrier()

    if rank == 0:
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        step_bytes = scene.algorithmic_bytes(n_total, cells_total)["step"]
        ms_per_step = ms / args.steps
        line = {"metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload,
                           "description": "%s: %d particles, cell 3, grid %dx%d, %d strips of cell columns (one per GPU), "
                                          "%s exchanged over NCCL every frame; the same scene at every GPU count" % (
                                              bench.WORKLOAD_TEXT[workload], n_total, gx, gy, world,
                                              "ghost tile columns" if tiles else "edge-column particles"),
                           "path": ("k_tile_frame per strip (edge tile columns first, ghost exchange on a second stream beside the interior)"
                                    if tiles else "k_phys + exchange + k_run_scan + k_rebin per strip"),
                           "tile_frames": st2["tile_frames"], "tile_fallbacks": st2["tile_fallbacks"],
                           "seed": hex(scene.SEED), "arith": "spv",
                           "l2": "per-GPU working set %.2f GB > 126 MB L2, no flush needed" % (n_local * 68 / 1e9),
                           "halo_bytes_per_step_per_rank": halo // max(1, st["steps_completed"]),
                           "verified": "after the timed frames, every rank: indices monotone, every particle in the slot range "
                                       "of its cell and in a column its strip owns, inside the world, |v|<=1; N conserved over all ranks",
                           "state_checksum": "%016x" % checksum_all, "frames_at_checksum": warm + args.steps},
                "clocks": clocks,
                "e2e": {"value": n_total * e2e_steps / e2e_dt, "unit": bench.UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                        "path": "per rank: write_slice x3 + write(settings) + step(1) + read_async x3 (capacity-sized) + one sync, pinned"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "per": "one GPU (max over ranks)",
                             "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                             "kernels": kernels,
                             "step": {"bytes": step_bytes, "gbs_all_gpus": step_bytes / ms_per_step / 1e6,
                                      "frac_of_n_x_peak": step_bytes / ms_per_step / 1e6 / (peak * world)}},
                "cpu_baseline": None, "scale_base": scale_base,
                "speedup_vs_1gpu_same_world": (scale_base["ms_per_step"] / ms_per_step
                                               if scale_base and "ms_per_step" in scale_base else None)}
        bench.emit(line)
    dist.destroy_process_group()
