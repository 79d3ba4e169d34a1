"""N > 1 host logic on CPU (gloo, world size 2): strip column split, per-strip packing by the C++
host mirror, the per-strip scene generator of bench_strips.py, and the max/sum-over-ranks plumbing.
No GPU, no NCCL: the device side of strips is covered by tests/test_gpu_strips.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import wrach_b200 as W
    from bench_strips import strip_scene
    from wrach_b200 import scene

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    dims, n = (300, 120), 24000
    wl = dict(n=n, dims=dims, pile=False)
    _, (gx, gy) = W.active_grid((0.0, 0.0, dims[0], dims[1]), 3)
    cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
    mine = strip_scene(scene, W, wl, rank, world, gx)
    # every particle of my scene sits inside my columns
    cx = np.floor(mine[:, 0] / np.float32(3)).astype(np.int64)
    assert cx.min() >= cols[0] and cx.max() < cols[1], (rank, cols, cx.min(), cx.max())
    # all ranks pack their own strip out of the union of all scenes
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    everything = np.concatenate(gathered)
    assert everything.shape[0] == n
    st = W.WrachState(W.WrachConfig(dims, cell_size=3), columns=cols)
    st.add_particles(everything)
    ind, pos, vel = st.create_packed_data()
    assert pos.shape[0] == mine.shape[0] == st.shader_settings.particles_in_frame_count
    assert tuple(st.shader_settings.grid_dimensions) == (gx, gy)  # the strip API keeps the GLOBAL grid
    # timing plumbing of bench_strips: max and sum over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    c = torch.tensor([float(pos.shape[0])], dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    assert int(c.item()) == n
    np.savez(os.path.join(out_dir, "strip%d.npz" % rank), ind=ind, pos=pos, vel=vel, cols=np.array(cols),
             everything=everything)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_strip_packing_equals_global_packing(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle as O

    world, port = 2, _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    strips = [np.load(os.path.join(tmp_path, "strip%d.npz" % r)) for r in range(world)]
    everything = strips[0]["everything"]
    dims = (300, 120)
    ow = O.OracleWorld(dims, 3)
    ow.add_particles(everything)
    gx, gy = ow.grid
    oind = ow.indices.astype(np.int64)
    covered = 0
    for s in strips:
        c0, c1 = s["cols"]
        width = c1 - c0
        ind = s["ind"].astype(np.int64)
        assert ind.shape[0] == width * gy + 2
        for cy in range(gy):
            for lx in (0, width // 2, width - 1):  # spot-check edge and middle cells of every row
                k_local, k_global = cy * width + lx, cy * gx + c0 + lx
                a, b = ind[k_local + 1], ind[k_local + 2]
                ga, gb = oind[k_global + 1], oind[k_global + 2]
                assert b - a == gb - ga
                assert np.array_equal(s["pos"][a:b], ow.positions_in[ga:gb])
                assert np.array_equal(s["vel"][a:b], ow.velocities_in[ga:gb])
        covered += int(ind[-1])
    assert covered == ow.n


def test_strip_scene_counts_and_ids():
    sys.path.insert(0, ROOT)
    import wrach_b200 as W
    from bench_strips import strip_scene
    from wrach_b200 import scene

    wl = dict(n=100001, dims=(655, 54), pile=False)
    _, (gx, gy) = W.active_grid((0.0, 0.0, 655.0, 54.0), 3)
    for world in (1, 2, 4, 8):
        parts = [strip_scene(scene, W, wl, r, world, gx) for r in range(world)]
        assert sum(p.shape[0] for p in parts) == wl["n"]
        for r, p in enumerate(parts):
            c0, c1 = W.PhysicsComputeWorker.strip_columns(gx, r, world)
            cx = np.floor(p[:, 0] / np.float32(3)).astype(np.int64)
            assert cx.min() >= c0 and cx.max() < c1
            assert p[:, 1].min() >= 0 and p[:, 1].max() < 54 and np.abs(p[:, 2:]).max() <= 0.5
