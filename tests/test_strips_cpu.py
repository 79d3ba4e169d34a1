"""N > 1 host logic on CPU (gloo, world size 2): strip column split, per-strip packing by the C++
host mirror, the column-bucketed global scene bench_strips.py feeds every rank (scene.generate_columns),
the self-checks and the additive checksum of a packed strip, and the max/sum-over-ranks plumbing.
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
    from wrach_b200 import scene

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    dims, n = (300, 120), 24000
    wl = dict(n=n, dims=dims, pile=False)
    _, (gx, gy) = W.active_grid((0.0, 0.0, dims[0], dims[1]), 3)
    cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
    mine = scene.generate_columns(wl["n"], dims[0], dims[1], cols, chunk=5000)
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
    # the collective re-bin's exchange plan: every rank contributes its row of
    # counts, all of them compute every segment offset from the gathered matrix, and what one rank
    # plans to send is what the other plans to receive
    from wrach_b200 import api
    rng = np.random.default_rng(11 + rank)
    my_row = np.concatenate([rng.integers(0, 1000, world), [5000 if rank == 0 else 100000]]).astype(np.int64)
    rows_t = [torch.zeros(world + 1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(rows_t, torch.from_numpy(my_row))
    rows = np.stack([r.numpy() for r in rows_t]).astype(np.uint32)
    send_off, recv_off, n_recv, over = api.strip_exchange_plan(rank, rows)
    assert over == -1 and n_recv == int(rows[:, rank].sum())
    assert list(send_off) == list(np.concatenate([[0], np.cumsum(rows[rank, :world])[:-1]]))
    assert list(recv_off) == list(np.concatenate([[0], np.cumsum(rows[:world, rank])[:-1]]))
    plans = [None] * world
    dist.all_gather_object(plans, (send_off.tolist(), recv_off.tolist(), n_recv))
    for s_ in range(world):       # segment s_ -> d_: the sender's extent fits before its next segment starts,
        for d_ in range(world):   # the receiver's likewise
            so, ro = plans[s_][0], plans[d_][1]
            cnt = int(rows[s_, d_])
            assert so[d_] + cnt <= (so[d_ + 1] if d_ + 1 < world else int(rows[s_, :world].sum()))
            assert ro[s_] + cnt <= (ro[s_ + 1] if s_ + 1 < world else plans[d_][2])
    tight = rows.copy()
    tight[0, world] = int(rows[:, 0].sum()) - 1   # strip 0 one slot short: EVERY rank must see it
    assert api.strip_exchange_plan(rank, tight)[3] == 0
    # the self-checks bench_strips runs on every rank's read-back, and the checksum it adds up
    assert scene.check_packed_invariants(ind, pos, vel, cols, gx, dims) == pos.shape[0]
    mine_sum = scene.state_checksum(ind, pos, vel, cols, gx)
    parts = torch.tensor([mine_sum & 0x7FFFFFFF, mine_sum >> 31], dtype=torch.int64)
    dist.all_reduce(parts, op=dist.ReduceOp.SUM)
    total = (int(parts[0]) + (int(parts[1]) << 31)) & 0xFFFFFFFFFFFFFFFF
    whole = W.WrachState(W.WrachConfig(dims, cell_size=3))
    whole.add_particles(scene.generate(n, dims[0], dims[1]))
    wi, wp, wv = whole.create_packed_data()
    assert total == scene.state_checksum(wi, wp, wv, (0, gx), gx), "the strips' checksums must add up to the whole world's"
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


def test_global_scene_bucketed_by_column():
    """Every GPU count sees the same scene: the strips' particles are a partition of generate()'s
    rows, in id order, whatever the number of strips and whatever the chunking."""
    sys.path.insert(0, ROOT)
    import wrach_b200 as W
    from wrach_b200 import scene

    n, dims = 100001, (655, 54)
    everything = scene.generate(n, dims[0], dims[1])
    _, (gx, gy) = W.active_grid((0.0, 0.0, 655.0, 54.0), 3)
    col = np.floor(everything[:, 0] / np.float32(3)).astype(np.int64)
    for world in (1, 2, 4, 8):
        parts = [scene.generate_columns(n, dims[0], dims[1], W.PhysicsComputeWorker.strip_columns(gx, r, world), chunk=30000)
                 for r in range(world)]
        assert sum(p.shape[0] for p in parts) == n
        for r, p in enumerate(parts):
            c0, c1 = W.PhysicsComputeWorker.strip_columns(gx, r, world)
            assert np.array_equal(p, everything[(col >= c0) & (col < c1)])


def test_packed_self_checks_catch_corruption():
    sys.path.insert(0, ROOT)
    import wrach_b200 as W
    from wrach_b200 import scene

    dims = (300, 120)
    st = W.WrachState(W.WrachConfig(dims, cell_size=3))
    st.add_particles(scene.generate(20000, dims[0], dims[1]))
    (gx, gy), _, _ = st.grid()
    ind, pos, vel = st.create_packed_data()
    assert scene.check_packed_invariants(ind, pos, vel, (0, gx), gx, dims) == 20000
    base = scene.state_checksum(ind, pos, vel, (0, gx), gx)
    for what, mutate in (("slot range", lambda i, p, v: p.__setitem__((7, 0), p[7, 0] + 3)),
                         ("outside", lambda i, p, v: p.__setitem__((7, 1), -1.0)),
                         ("|v|", lambda i, p, v: v.__setitem__((7, 1), 1.5)),
                         ("monotone", lambda i, p, v: i.__setitem__(5, i[6] + 1))):
        i2, p2, v2 = ind.copy(), pos.copy(), vel.copy()
        mutate(i2, p2, v2)
        with pytest.raises(AssertionError):
            scene.check_packed_invariants(i2, p2, v2, (0, gx), gx, dims)
    p2 = pos.copy()
    a = int(np.flatnonzero(np.diff(ind[1:].astype(np.int64)) >= 2)[0])  # a cell with two particles: swap them
    s0 = int(ind[a + 1])
    p2[[s0, s0 + 1]] = p2[[s0 + 1, s0]]
    assert scene.state_checksum(ind, p2, vel, (0, gx), gx) != base, "the checksum must see the order inside a cell"
