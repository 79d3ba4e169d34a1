#!/usr/bin/env python3
"""Multi-process check of the NCCL strip path (run under torchrun, one rank per GPU): every rank
steps its strip for a number of frames and compares each of its cells, bit for bit and in canonical
order, with the single-device CPU oracle run on the whole world.  Exits non-zero on any mismatch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/tools/strip_nccl_check.py --case uniform|pile|far|halo|256m

Cases: `uniform` (a 330 k world, 40 frames in two batches), `pile` (skewed occupancy: dense runs on
both sides of every strip boundary), `far` (wild first-frame velocities: particles leaving for
non-adjacent strips, the cross-rank generic re-bin), `halo` (the opt-in 3x3 neighbour mode: ghost
columns exchanged every frame), `tiles-far` / `tiles-crowd` (strips on tile frames that meet a frame the
tiles cannot hold -- right after the upload / in the middle of a batch: the vote, the checkpoint replay), `256m` (2 frames of BASELINE.json configs[4] at full size: rank 0
runs the OpenMP oracle once and shares the result through /dev/shm).
tests/test_gpu_multirank.py launches this over every GPU count the box offers."""
import argparse
import os
import shutil
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wrach_b200 as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.test_gpu_strips import assert_strips_equal_oracle  # noqa: E402
from wrach_b200 import scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="uniform", choices=["uniform", "pile", "far", "halo", "tiles-far", "tiles-crowd", "256m"])
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
threads = max(1, len(os.sched_getaffinity(0)) // world)

CASES = {
    # dims, particles, frames (two batches), pile?, capacity factor over N, velocity scale, neighbours
    "uniform": dict(dims=(900, 500), n=330000, frames=40, pile=False, vscale=1.0, nb=False),
    "pile": dict(dims=(720, 300), n=200000, frames=10, pile=True, vscale=1.0, nb=False),
    "far": dict(dims=(900, 300), n=150000, frames=6, pile=False, vscale=60.0, nb=False),
    "halo": dict(dims=(900, 300), n=200000, frames=12, pile=False, vscale=1.0, nb=True),
    # wide enough for two tile columns per strip at 8 ranks: the strips run on tile frames
    "tiles-far": dict(dims=(2200, 300), n=400000, frames=24, pile=False, vscale=1.0, nb=False, wild_every=17),
    "tiles-crowd": dict(dims=(2200, 300), n=400000, frames=12, pile=False, vscale=1.0, nb=False, crowd=True),
}


def unique_id():
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(W.PhysicsComputeWorker.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    return bytes(uid.cpu().numpy())


def small_case(c):
    dims, n, frames = c["dims"], c["n"], c["frames"]
    p = O.generate_scene(n, dims[0], dims[1], seed=2024 + len(args.case), pile=c["pile"])
    if c["vscale"] != 1.0:
        p[:, 2:] *= np.float32(c["vscale"])  # several cells (and, on narrow strips, several strips) per frame
    if c.get("wild_every"):
        p[::c["wild_every"], 2:] *= np.float32(400.0)  # a first frame the tiles cannot hold: all strips fall back together
    if c.get("crowd"):
        # a cell passes 255 particles at frame 3 or 4, in the middle of the second batch: every strip goes
        # back to the checkpoint (the packed copy of the last read-back), replays, and leaves the tiles
        rng = np.random.default_rng(6)
        still = np.zeros((140, 4), np.float32)
        still[:, 0] = 150.1 + rng.random(140, dtype=np.float32) * np.float32(2.8)
        still[:, 1] = 99.1 + rng.random(140, dtype=np.float32) * np.float32(2.8)
        movers = np.zeros((140, 4), np.float32)
        movers[:, 0] = np.float32(155.5)
        movers[:, 1] = 99.1 + rng.random(140, dtype=np.float32) * np.float32(2.8)
        movers[:, 2] = np.float32(-1.0)
        p = np.concatenate([p, still, movers])
        n = p.shape[0]
    config = W.WrachConfig(dims, cell_size=3)
    full = W.WrachState(config)
    (gx, gy), _, cap = full.grid()
    cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
    st = W.WrachState(config, columns=cols)
    st.add_particles(p)
    g = full.shader_settings.copy()
    g.particles_in_frame_count = 0
    w = W.PhysicsComputeWorker(g, 0, max(cap, n), device=local, strip=(rank, world, unique_id()))
    if c["nb"]:
        w.set_neighbour_mode(True)
    W.maybe_upload_to_gpu(w, st)
    ow = O.OracleWorld(dims, 3, capacity=max(cap, 2 * n), neighbours=c["nb"])
    ow.add_particles(p)
    done = 0
    for upto in (1, frames // 2, frames):  # compare after the first frame and after each batch
        ow.step(upto - done, threads=threads)
        w.step(upto - done)
        done = upto
        assert_strips_equal_oracle([w], [cols], (gx, gy), ow, "case %s rank %d after %d frames" % (args.case, rank, upto))
    st_ = w.stats()
    if args.case == "far":
        assert st_["slow_path_steps"] >= 1, st_
    if args.case.startswith("tiles-"):
        assert st_["tile_fallbacks"] >= 1 and st_["tile_frames"] >= 1, st_
    print("rank %d/%d case %s: columns %s bit-exact vs oracle after %d frames; halo bytes %d, slow-path frames %d, tile frames %d, fallbacks %d" % (
        rank, world, args.case, cols, frames, st_["halo_bytes_sent"], st_["slow_path_steps"], st_["tile_frames"], st_["tile_fallbacks"]), flush=True)
    w.close()


def full_size_case():
    wl = scene.WORKLOADS["256m"]
    n, dims, frames = wl["n"], wl["dims"], 2
    config = W.WrachConfig(dims, cell_size=3)
    _, (gx, gy) = W.active_grid((0.0, 0.0, dims[0], dims[1]), 3)
    cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
    st = W.WrachState(config, columns=cols)
    st.add_particles(scene.generate_columns(n, dims[0], dims[1], cols))
    g = st.shader_settings.copy()
    n_local = g.particles_in_frame_count
    g.particles_in_frame_count = 0
    _, _, cap = st.grid()
    w = W.PhysicsComputeWorker(g, 0, max(cap, int(n_local * 1.1)), device=local, strip=(rank, world, unique_id()))
    W.maybe_upload_to_gpu(w, st)
    st.close()
    w.step(frames)
    shm = "/dev/shm/wrach_multirank_%s" % os.environ.get("MASTER_PORT", "0")
    if rank == 0:
        os.makedirs(shm, exist_ok=True)
        ow = O.OracleWorld(dims, 3)
        ow.add_particles(scene.generate_fast(n, dims[0], dims[1]))
        ow._store = None
        ow.step(frames, threads=0)
        np.save(os.path.join(shm, "ind.npy"), ow.indices)
        np.save(os.path.join(shm, "pos.npy"), ow.positions_in[:ow.n])
        np.save(os.path.join(shm, "vel.npy"), ow.velocities_in[:ow.n])
        del ow
    dist.barrier()
    try:
        ow = types.SimpleNamespace(indices=np.load(os.path.join(shm, "ind.npy"), mmap_mode="r"),
                                   positions_in=np.load(os.path.join(shm, "pos.npy"), mmap_mode="r"),
                                   velocities_in=np.load(os.path.join(shm, "vel.npy"), mmap_mode="r"))
        ow.n = int(ow.indices[-1])
        assert_strips_equal_oracle([w], [cols], (gx, gy), ow, "256m rank %d after %d frames" % (rank, frames))
        total = torch.tensor([int(w.read_vec(W.Buffers.INDICES_MAIN)[-1])], dtype=torch.int64, device="cuda")
        dist.all_reduce(total)
        assert int(total.item()) == n, "%d particles over all strips, scene has %d" % (int(total.item()), n)
        print("rank %d/%d case 256m: columns %s bit-exact vs the single-device oracle after %d frames" % (
            rank, world, cols, frames), flush=True)
    finally:
        dist.barrier()
        if rank == 0:
            shutil.rmtree(shm, ignore_errors=True)
    w.close()


if args.case == "256m":
    full_size_case()
else:
    small_case(CASES[args.case])
dist.barrier()
dist.destroy_process_group()
