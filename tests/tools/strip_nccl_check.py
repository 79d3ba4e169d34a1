#!/usr/bin/env python3
"""Multi-process check of the NCCL strip path (run under torchrun, one rank per GPU): every rank
steps its strip for a number of frames and compares each of its cells with the single-device CPU
oracle run on the whole world.  Prints one line per rank; exits non-zero on any mismatch.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/tools/strip_nccl_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wrach_b200 as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.test_gpu_strips import assert_strips_equal_oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims, n, frames = (900, 500), 330000, 40
p = O.generate_scene(n, dims[0], dims[1], seed=2024)
config = W.WrachConfig(dims, cell_size=3)
full = W.WrachState(config)
(gx, gy), _, cap = full.grid()
cols = W.PhysicsComputeWorker.strip_columns(gx, rank, world)
st = W.WrachState(config, columns=cols)
st.add_particles(p)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(W.PhysicsComputeWorker.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
g = full.shader_settings.copy()
g.particles_in_frame_count = 0
w = W.PhysicsComputeWorker(g, 0, cap, device=local, strip=(rank, world, bytes(uid.cpu().numpy())))
W.maybe_upload_to_gpu(w, st)
ow = O.OracleWorld(dims, 3)
ow.add_particles(p)
ow.step(frames, threads=4)
w.step(frames // 2)
w.step(frames - frames // 2)
assert_strips_equal_oracle([w], [cols], (gx, gy), ow, "rank %d after %d frames" % (rank, frames))
print("rank %d/%d: strip columns %s bit-exact vs oracle after %d frames, halo bytes sent %d" % (
    rank, world, cols, frames, w.stats()["halo_bytes_sent"]), flush=True)
w.close()
dist.barrier()
dist.destroy_process_group()
