"""The C++ host mirror (SpatialBin / ParticleStore / WrachState, include/wrach_host.h) against the
reference's own host-side unit tests and against the oracle's packer.  CPU only."""
import ctypes
import os
import re

import numpy as np

import wrach_b200 as W
from oracle import oracle as O

f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_header_symbols_exported():
    from wrach_b200 import _ffi, api
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "wrach_host.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(wrach_(?:host|state|plugin|api|config)_[a-z_0-9]+)\s*\(", text)))
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(api.HOST_SYMBOLS)


def test_cell_coord_kats(kats):
    for k in kats["cell_coord"]:
        got = [W.get_cell_coord(k["position"][0], k["cell_size"]), W.get_cell_coord(k["position"][1], k["cell_size"])]
        assert got == k["coord"], k["cite"]


def test_active_cells_kats(kats):
    for k in kats["active_cells"]:
        cells, grid = W.get_active_cells(k["viewport"], k["cell_size"])
        assert [list(c) for c in cells] == k["cells"], k["cite"]
        if k["grid"] is not None:
            assert list(grid) == k["grid"]


def test_packed_data_and_capacity_kats(kats):
    for k in kats["packed_data"]:
        st = W.WrachState(W.WrachConfig((int(k["viewport"][2]), int(k["viewport"][3])), cell_size=k["cell_size"]))
        st.add_particles(np.array(k["particles"], f32))
        ind, pos, vel = st.create_packed_data()
        assert ind.tolist() == k["indices"], k["cite"]
        assert np.array_equal(pos, np.array(k["positions"], f32)) and np.array_equal(vel, np.array(k["velocities"], f32))
        assert st.gpu_uploads_pending == 2  # PackedData + Settings (state.rs:95-100)
        assert st.shader_settings.particles_in_frame_count == len(k["positions"])
    for k in kats["capacity"]:
        st = W.WrachState(W.WrachConfig((int(k["viewport"][2]), int(k["viewport"][3])), cell_size=k["cell_size"]))
        assert st.grid()[2] == k["max_particles_per_frame"], k["cite"]


def test_default_config_matches_reference():
    c = W.WrachConfig()
    assert (c.dimensions, c.cell_size, c.boundaries_as_dimensions) == ((480, 352), 3, False)  # config_app.rs:24-36
    st = W.WrachState(c)
    (gx, gy), total, cap = st.grid()
    assert (gx, gy) == (161, 118) and total == 161 * 118 + 2 and cap == 188082


def test_packer_equals_oracle_on_random_scene_incrementally():
    dims, cell = (300, 200), 3
    st = W.WrachState(W.WrachConfig(dims, cell_size=cell))
    ow = O.OracleWorld(dims, cell, capacity=60000)
    rng = np.random.default_rng(0)
    for chunk in range(3):  # add_particles re-packs the whole store every time (state.rs:90-101)
        p = O.generate_scene(15000, dims[0] * 1.1, dims[1] * 1.1, seed=chunk)  # some land off-viewport
        p[:5, :2] = -p[:5, :2]
        st.add_particles(p)
        ow.add_particles(p)
        ind, pos, vel = st.create_packed_data()
        n = ow.n
        assert pos.shape[0] == n
        assert np.array_equal(ind, ow.indices)
        assert np.array_equal(pos, ow.positions_in[:n]) and np.array_equal(vel, ow.velocities_in[:n])


def test_distance_threshold_equivalence():
    """k_phys tests `sqrt_rn(d2) > 1` as `d2 > 0x3F800001`: check every float around 1 and the
    special values (the GPU sqrt is IEEE round-to-nearest like numpy's)."""
    t = np.array([0x3F800001], np.uint32).view(f32)[0]
    bits = np.arange(0x3F800000 - 4096, 0x3F800000 + 4096, dtype=np.uint32)
    d2 = bits.view(f32)
    assert np.array_equal(np.sqrt(d2) > f32(1.0), d2 > t)
    special = np.array([0.0, 1e-45, 0.25, 0.99999994, 1.0, 1.0000001, 1.0000002, 4.0, np.inf, np.nan], f32)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(np.sqrt(special) > f32(1.0), special > t)
    rng = np.random.default_rng(1)
    r = (rng.random(1_000_000, dtype=f32) * f32(4.0)).astype(f32)
    assert np.array_equal(np.sqrt(r) > f32(1.0), r > t)


def test_fast_key_equals_divide_key_sweep():
    """k_phys / k_tile_frame classify a move with exact compares against the bounds of the source cell
    instead of the reference's  u32(floor((x - anchor) / f32(cell_size)))  (particles_per_cell.wgsl:14-27).
    That rests on  floor(fl(rel / cs)) == floor(rel / cs)  for an integer cell size and 0 <= rel < 2^23
    (wrach_kernels.cuh: finish_particle; the bound is enforced by validate_settings): the rounded
    quotient of the float just below a multiple k*cs must not reach k.  Swept here for every multiple
    below 2^23, the floats around it, and a set of cell sizes incl. the reference's 3."""
    lim = 1 << 23
    for cs in list(range(1, 14)) + [100, 255, 4096, 65535]:
        k = np.arange(1, lim // cs + 1, dtype=np.int64)
        m = (k * cs).astype(f32)                       # exact: integers below 2^24
        assert np.array_equal(m.astype(np.int64), k * cs)
        c = f32(cs)
        lo = m
        for step in range(1, 4):                       # 1, 2, 3 floats below / above the multiple
            lo = np.nextafter(lo, f32(-np.inf), dtype=f32)
            exact = np.floor(lo.astype(np.float64) / float(cs)).astype(np.int64)  # (k - 1, or k - 2 once 3 floats span a cell)
            assert np.all(exact < k) and np.array_equal(np.floor(lo / c).astype(np.int64), exact), (cs, -step)
        hi = m
        for step in range(0, 3):
            exact = np.floor(hi.astype(np.float64) / float(cs)).astype(np.int64)
            assert np.all(exact >= k) and np.array_equal(np.floor(hi / c).astype(np.int64), exact), (cs, step)
            hi = np.nextafter(hi, f32(np.inf), dtype=f32)
    # ... and on random values (the compare-based code is: new cell = old + (rel >= hi) - (rel < lo))
    rng = np.random.default_rng(7)
    for cs in (3, 5, 7, 13):
        rel = (rng.random(2_000_000, dtype=f32) * f32(lim - 1)).astype(f32)
        key = np.floor(rel / f32(cs)).astype(np.int64)
        lo_b, hi_b = (key * cs).astype(f32), ((key + 1) * cs).astype(f32)
        assert np.all((rel >= lo_b) & (rel < hi_b)), cs


def test_scene_generator_equals_oracle_generator():
    from wrach_b200 import scene
    for pile in (False, True):
        a = scene.generate(50000, 1366.0, 1024.0, first_id=12345, pile=pile, chunk=7777)
        b = O.generate_scene(50000, 1366.0, 1024.0, seed=scene.SEED, first_id=12345, pile=pile)
        assert np.array_equal(a, b)
    assert a[:, 0].min() >= 0 and a[:, 0].max() < 1366 and np.abs(a[:, 2:]).max() <= 0.5
    c = scene.generate_fast(50000, 1366.0, 1024.0, first_id=12345, pile=True)
    assert np.array_equal(a, c)


def test_baseline_config_geometry():
    """SURVEY.md §8d / BASELINE.md §3 table: grids, cell counts and capacities of the five configs."""
    from wrach_b200 import scene
    expect = {"1m-scene": ((494, 351), 173394, 1716606), "1m": ((456, 342), 155952, 1543928),
              "16m": ((1822, 1366), 2488852, 24639638), "64m-pile": ((3643, 2731), 9949033, None),
              "256m": ((21845, 1821), 39779745, None)}
    for name, (grid, cells, cap) in expect.items():
        wl = scene.WORKLOADS[name]
        _, g = W.active_grid((0.0, 0.0, wl["dims"][0], wl["dims"][1]), 3)
        assert g == grid and g[0] * g[1] == cells
        if cap:
            assert W.max_particles_per_frame(cells, 3) == cap
    assert scene.algorithmic_bytes(1 << 24, 2488852)["step"] == 64 * (1 << 24) + 16 * 2488852


def test_viewport_streaming_host_side():
    """update_from_gpu + set_viewport with the oracle standing in for the device:
    the C++ store must hand the next window exactly what the numpy restatement hands it."""
    from tests.util import WindowedOracle
    rng = np.random.default_rng(7)
    n = 30000
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rng.uniform(0, 540, n)
    p[:, 1] = rng.uniform(0, 240, n)
    p[:, 2:] = rng.uniform(-0.5, 0.5, (n, 2))
    state = W.WrachState(W.WrachConfig((240, 240), cell_size=3))
    state.add_particles(p)
    assert state.stored_particles == n
    ref = WindowedOracle(p, (0, 0, 240, 240))
    for viewport in ((150, 0, 390, 240), (300, 0, 540, 240), (0, 0, 240, 240)):
        ind, pos, vel = state.create_packed_data()
        assert np.array_equal(ind, ref.indices) and np.array_equal(pos, ref.pos[:ref.n]) and np.array_equal(vel, ref.vel[:ref.n])
        ref.step(3)
        state.set_packed_data(ref.indices, ref.pos[:ref.n], ref.vel[:ref.n])  # what tick would have read back
        state.update_from_gpu()
        ref.update_from_gpu()
        assert state.stored_particles == n
        state.set_viewport(viewport)
        ref.set_viewport(viewport)
        assert state.gpu_uploads_pending >= 2
        s = state.shader_settings
        assert list(s.view_anchor) == [float(viewport[0]), float(viewport[1])] and s.particles_in_frame_count == ref.n
    ind, pos, vel = state.create_packed_data()
    assert np.array_equal(ind, ref.indices) and np.array_equal(pos, ref.pos[:ref.n]) and np.array_equal(vel, ref.vel[:ref.n])
    # the window must keep its grid and sit on a cell boundary
    for bad in ((1, 0, 241, 240), (0, 0, 300, 240), (0, 3, 240, 240)):
        try:
            state.set_viewport(bad)
        except W.WrachCudaError as e:
            assert e.status == -1
        else:
            raise AssertionError("set_viewport%r accepted" % (bad,))
    # a read-back with another layout is refused
    state.set_packed_data(np.zeros(5, np.uint32), np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    try:
        state.update_from_gpu()
    except W.WrachCudaError as e:
        assert e.status == -5
    else:
        raise AssertionError("update_from_gpu accepted a foreign layout")


def test_large_host_arrays_fall_back_to_heap_memory_without_a_device():
    """PackedData arrays of 32 MB and more come from cudaMallocHost (wrach_host.hpp: HostArray); on a box
    without a CUDA device the allocation and the matching free must quietly use the heap."""
    import wrach_b200 as W
    from wrach_b200 import scene
    n, dims = 4_500_000, (3000, 2000)
    st = W.WrachState(W.WrachConfig(dims, cell_size=3))
    st.add_particles(scene.generate_fast(n, dims[0], dims[1]))
    ind, pos, vel = st.create_packed_data()          # three arrays, positions / velocities 36 MB each
    assert pos.shape == (n, 2) and int(ind[-1]) == n and pos.nbytes >= 32 << 20
    st.set_packed_data(ind, pos, vel)                # ... and the state's own copies
    _, p2, v2 = st.packed_data
    assert np.array_equal(p2[:n], pos) and np.array_equal(v2[:n], vel)
    st.close()

