"""get_buffer for a renderer (runners/bevy/src/plugin/bind_groups.rs:61-83): the
CUDA half of CUDA <-> Vulkan external-memory interop.  POSITIONS_IN is moved into a shareable allocation
and exported as an opaque POSIX file descriptor (what VK_KHR_external_memory_fd imports).  No Vulkan
loader exists in this image, so the other half is played by CUDA itself: the descriptor is imported
and mapped a second time (cuMemImportFromShareableHandle) and must show, byte for byte, the packed
frame the oracle holds -- before and after more frames have run through the moved buffer."""
import ctypes
import os

import numpy as np
import pytest

import wrach_b200 as W
from oracle import oracle as O
from tests.util import assert_same_state, f32, make_pair
from wrach_b200 import Buffers, _ffi

pytestmark = pytest.mark.gpu


def read_through_fd(fd, alloc_bytes, nbytes):
    out = np.empty(nbytes, np.uint8)
    rc = _ffi.lib().wrach_cuda_selftest_import_fd(0, fd, alloc_bytes, out.ctypes.data, nbytes)
    assert rc == 0, _ffi.lib().wrach_cuda_last_error(None)
    return out


@pytest.mark.parametrize("tiles", ["1", "0"])
def test_exported_positions_show_the_packed_frame(tiles, monkeypatch):
    monkeypatch.setenv("WRACH_TILES", tiles)
    dims, n = (420, 300), 60000
    p = O.generate_scene(n, dims[0], dims[1], seed=9)
    ow, w = make_pair(dims, 3, p)
    ow.step(3)
    w.step(3)
    fd, alloc = w.export_buffer_fd(Buffers.POSITIONS_IN)   # moves the buffer; frames keep running through it
    assert fd >= 0 and alloc >= ow.capacity * 8 and alloc % 4096 == 0
    w.settle()
    got = read_through_fd(fd, alloc, ow.n * 8).view(np.float32).reshape(-1, 2)
    assert np.array_equal(got.view(np.uint32), ow.positions_in[:ow.n].view(np.uint32))
    assert_same_state(ow, w, "after the move")
    ow.step(5)
    w.step(5)
    w.settle()                                             # the renderer's per-frame call
    fd2, alloc2 = w.export_buffer_fd(Buffers.POSITIONS_IN)  # a second descriptor of the same allocation
    assert alloc2 == alloc
    got = read_through_fd(fd2, alloc2, ow.n * 8).view(np.float32).reshape(-1, 2)
    assert np.array_equal(got.view(np.uint32), ow.positions_in[:ow.n].view(np.uint32))
    st = w.stats()
    assert (st["tile_frames"] > 0) == (tiles == "1")
    fdv, allocv = w.export_buffer_fd(Buffers.VELOCITIES_IN)
    got = read_through_fd(fdv, allocv, ow.n * 8).view(np.float32).reshape(-1, 2)
    assert np.array_equal(got.view(np.uint32), ow.velocities_in[:ow.n].view(np.uint32))
    assert_same_state(ow, w, "8 frames")
    w.close()


def test_only_what_a_renderer_binds_can_be_exported():
    dims = (60, 40)
    ow, w = make_pair(dims, 3, O.generate_scene(500, dims[0], dims[1], seed=1))
    with pytest.raises(W.WrachCudaError) as e:
        w.export_buffer_fd(Buffers.INDICES_MAIN)
    assert e.value.status == -1
    w.close()
