"""The NCCL strip path under torchrun, one process per GPU, over every GPU count the box offers
(2, 4, 8): each rank holds its cells bit-exact to the single-device OpenMP oracle
(tests/tools/strip_nccl_check.py).  The reference checks its GPU passes against its CPU packing the
same way (runners/bevy/src/compute/03_prefix_sum.rs:151-260).  Skips on a single-GPU box, where
tests/test_gpu_strips.py covers the same kernels with in-process strips."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_case(n_ranks, case, timeout):
    env = dict(os.environ)
    env.pop("WRACH_PDL", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_ranks),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "tools", "strip_nccl_check.py"), "--case", case]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, "torchrun x%d case %s failed:\n%s\n%s" % (n_ranks, case, r.stdout[-3000:], r.stderr[-6000:])
    assert r.stdout.count("bit-exact") == n_ranks, r.stdout[-3000:]


@pytest.mark.parametrize("case", ["uniform", "pile", "far", "halo", "tiles-far", "tiles-crowd"])
@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_nccl_strips_equal_single_device_oracle(n_ranks, case):
    if gpu_count() < n_ranks:
        pytest.skip("needs %d GPUs" % n_ranks)
    run_case(n_ranks, case, timeout=600)


def test_nccl_strips_256m_world_two_frames():
    """BASELINE.json configs[4] at full size over all the GPUs of the box, 2 frames."""
    n = gpu_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    run_case(8 if n >= 8 else 4 if n >= 4 else 2, "256m", timeout=1500)
