"""Parity proper: the CUDA path, called through the C ABI, against the CPU oracle on identical
seeded inputs.  Bit-exact: cell keys / indices / packed (canonical, stable) order, positions and
velocities (0 ULP, same arithmetic variant), after 1..N frames."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_same_state, f32, make_pair, read_state

pytestmark = pytest.mark.gpu
ARITHS = [O.ARITH_UNFUSED, O.ARITH_SPV]


# ---- the reference's own GPU tests, replayed through the C ABI --------------------------------

def test_ref_prefix_sum_small_and_large(kats):
    """03_prefix_sum.rs:151-260: device indices == create_packed_data().indices after 4 ticks."""
    for k in kats["gpu_equals_cpu_indices"]:
        p = np.array(k["particles"], f32)
        ow, w = make_pair(k["dimensions"], k["cell_size"], p)
        cpu_indices, _, _ = ow.pack(p)
        for _ in range(k["ticks"]):
            w.step(1)
            ind, _, _ = read_state(w)
        assert np.array_equal(ind, cpu_indices), k["cite"]
        if k["indices"] is not None:
            assert ind.tolist() == k["indices"]


def test_ref_packed_data(kats):
    """04_pack_particle_data.rs:73-140."""
    k = kats["packed_positions_after_ticks"]
    ow, w = make_pair(k["dimensions"], k["cell_size"], np.array(k["particles"], f32))
    for _ in range(k["ticks"]):
        w.step(1)
    _, pos, _ = read_state(w)
    assert np.array_equal(pos[:4], np.array(k["positions"], f32))


def test_ref_api_smoke(kats):
    """runners/api/src/lib.rs:102-126: three coincident particles, 5 ticks, capacity-sized read-back."""
    k = kats["api_smoke"]
    ow, w = make_pair(k["dimensions"], k["cell_size"], np.array(k["particles"], f32))
    for _ in range(k["ticks"]):
        w.step(1)
        ow.step(1)
    ind, pos, vel = read_state(w)
    assert pos.shape[0] == k["readback_len"] == vel.shape[0]
    assert tuple(pos[0]) != (0.0, 0.0) and tuple(vel[0]) != (0.0, 0.0)
    assert_same_state(ow, w)


# ---- seeded scenes vs the oracle --------------------------------------------------------------

@pytest.mark.parametrize("arith", ARITHS)
@pytest.mark.parametrize("dims,cell,n", [((10, 10), 3, 40), ((64, 48), 3, 2304), ((333, 217), 3, 54000),
                                          ((500, 300), 6, 100000), ((97, 61), 1, 3000), ((120, 90), 5, 9000)])
def test_uniform_scene_every_step(arith, dims, cell, n):
    p = O.generate_scene(n, dims[0], dims[1], seed=1234 + n)
    ow, w = make_pair(dims, cell, p, arith=arith, capacity=2 * n + 64)
    assert_same_state(ow, w, "upload")
    for t in range(12):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))


@pytest.mark.parametrize("arith", ARITHS)
def test_batched_steps_equal_single_steps(arith):
    n, dims = 200000, (640, 420)
    p = O.generate_scene(n, dims[0], dims[1], seed=77)
    ow, w = make_pair(dims, 3, p, arith=arith)
    ow.step(25, threads=4)
    w.step(10)
    w.step(15)  # enqueued back to back, no read in between
    assert_same_state(ow, w, "25 steps")
    assert w.stats()["slow_path_steps"] == 0


def test_wild_first_frame_velocities_take_the_generic_path():
    """|v| far above the cell size: particles jump anywhere on frame 1 (clamped after, particles.rs:103-104)."""
    n, dims = 30000, (300, 200)
    p = O.generate_scene(n, dims[0], dims[1], seed=5)
    p[:, 2:] *= f32(500.0)
    ow, w = make_pair(dims, 3, p)
    for t in range(5):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))
    assert w.stats()["slow_path_steps"] >= 1


def test_far_mover_in_the_middle_of_a_batch():
    """A batch whose 1st frame needs the generic path: the frames behind it are replayed."""
    n, dims = 20000, (240, 160)
    p = O.generate_scene(n, dims[0], dims[1], seed=9)
    p[::7, 2:] *= f32(300.0)
    ow, w = make_pair(dims, 3, p)
    ow.step(8)
    w.step(8)
    assert_same_state(ow, w, "8 steps")
    st = w.stats()
    assert st["slow_path_steps"] >= 1 and st["steps_completed"] == 8


@pytest.mark.parametrize("arith", ARITHS)
def test_pile_skewed_occupancy(arith):
    """config 3 in miniature: y = H*u^4 -> bottom rows hold hundreds of particles per cell."""
    n, dims = 120000, (300, 400)
    p = O.generate_scene(n, dims[0], dims[1], seed=11, pile=True)
    ow, w = make_pair(dims, 3, p, arith=arith, capacity=2 * n)
    for t in range(6):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))
    # frame 1 walks the dense runs inside k_rebin and raises the flag; from frame 2 on the frame is
    # k_phys + k_run_scan + k_rebin + k_rebin_dense
    # (with the fused tile frames on, the very first launch pair -- unpack + tile frame -- finds the
    # scene too dense for the tiles and the worker falls back for good: two launches more)
    st = w.stats()
    assert st["kernel_launches"] == 3 + 5 * 4 + (2 if st["tile_fallbacks"] else 0)
    assert st["tile_frames"] == 0
    ow.step(5)
    w.step(5)
    assert_same_state(ow, w, "batch of 5 more")


def test_every_run_over_full():
    """Twice the staging capacity everywhere: every run of 256 cells takes the dense physics mode and
    the general re-bin path; > 2048 listed source ranges, i.e. more than one batch of k_rebin_dense."""
    dims = (1600, 1100)
    n = 18 * 534 * 367
    p = O.generate_scene(n, dims[0], dims[1], seed=23)
    ow, w = make_pair(dims, 3, p, capacity=n + 1024)
    for t in range(3):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))
    ow.step(3)
    w.step(3)
    assert_same_state(ow, w, "batch of 3 more")
    assert w.stats()["slow_path_steps"] == 0


def test_dense_band_next_to_normal_cells():
    """A few very heavy rows (thousands per cell) in an otherwise normal world: staged runs, dense
    runs and the rows between them (listed row changers unknown on one side) in one frame."""
    dims = (900, 300)
    rng = np.random.default_rng(5)
    base = O.generate_scene(150000, dims[0], dims[1], seed=3)
    heavy = np.empty((400000, 4), f32)
    heavy[:, 0] = rng.uniform(0, dims[0], len(heavy))
    heavy[:, 1] = rng.uniform(148.5, 153.5, len(heavy))  # rows 49..51
    heavy[:, 2:] = rng.uniform(-0.5, 0.5, (len(heavy), 2))
    p = np.concatenate([base, heavy.astype(f32)])
    ow, w = make_pair(dims, 3, p, capacity=len(p) + 1024)
    for t in range(4):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))


def test_everything_in_one_cell():
    n = 5000
    p = np.zeros((n, 4), f32)
    p[:, 0] = 4.0 + (np.arange(n) % 97) * f32(0.01)
    p[:, 1] = 4.0 + (np.arange(n) % 89) * f32(0.01)
    p[:, 2] = f32(0.3)
    ow, w = make_pair((30, 30), 3, p, capacity=n + 10)
    for t in range(4):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))


def test_boundaries_corners_and_nan():
    p = np.array([[9.5, 5.0, 3.0, 0.0], [0.2, 0.1, -0.5, -0.5], [10.0, 10.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0],
                  [10.0, 0.0, 0.9, -0.9], [5.0, 5.0, np.nan, 0.0], [5.1, 5.1, 0.0, np.inf]], f32)
    ow, w = make_pair((10, 10), 5, p)
    for t in range(4):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))


def test_empty_world_and_ragged_tiles():
    ow, w = make_pair((10, 10), 3, np.zeros((0, 4), f32))
    w.step(3)
    ow.step(3)
    assert_same_state(ow, w)
    # grid 129 x 3 = 387 cells: not a multiple of any tile size, particles only in the last cell
    p = np.array([[386.9, 8.9, 0.0, 0.0]] * 3, f32)
    ow, w = make_pair((386, 8), 3, p)
    ow.step(2)
    w.step(2)
    assert_same_state(ow, w)


def test_capacity_and_argument_errors():
    import wrach_b200
    from wrach_b200 import Buffers
    ow, w = make_pair((10, 10), 3, np.zeros((0, 4), f32))
    with pytest.raises(wrach_b200.WrachCudaError) as e:
        w.write_slice(Buffers.POSITIONS_IN, np.zeros((ow.capacity + 1, 2), f32))
    assert e.value.status == -2
    bad = wrach_b200.WorldSettings()
    bad.view_dimensions[:] = [100.0, 100.0]
    bad.grid_dimensions[:] = [4, 4]  # does not cover a 100x100 viewport at cell 3
    bad.cell_size = 3
    with pytest.raises(wrach_b200.WrachCudaError) as e:
        w.write(Buffers.WORLD_SETTINGS_UNIFORM, bad)
    assert e.value.status == -1


# ---- BASELINE.json configs[2] and [3] at full size: bit-exact against the (OpenMP) oracle ---------

def check_invariants(ow, w, n, dims):
    """Size-independent properties of a packed frame, on the GPU's own read-back."""
    ind, pos, vel = read_state(w)
    assert ind[0] == 0 and ind[-1] == n and np.all(np.diff(ind.astype(np.int64)) >= 0)  # monotone, conserves N
    p, v = pos[:n], vel[:n]
    assert np.all((p[:, 0] >= 0) & (p[:, 0] <= dims[0]) & (p[:, 1] >= 0) & (p[:, 1] <= dims[1]))
    assert np.all(np.abs(v) <= 1.0)
    # every particle sits in the slot range of the cell its position keys to (sortedness)
    gx = ow.grid[0]
    key = (np.floor(p[:, 1] / f32(3)).astype(np.int64) * gx + np.floor(p[:, 0] / f32(3)).astype(np.int64))
    slot = np.arange(n, dtype=np.int64)
    assert np.all((ind[key + 1] <= slot) & (slot < ind[key + 2]))


def test_config2_sixteen_million_full_size():
    n, dims = 1 << 24, (5464, 4096)
    p = O.generate_scene(n, dims[0], dims[1])
    ow, w = make_pair(dims, 3, p)
    assert (ow.grid, ow.cells, ow.capacity) == ((1822, 1366), 2488852, 24639638)
    done = 0
    for upto in (1, 6):
        ow.step(upto - done, threads=0)
        w.step(upto - done)
        done = upto
        assert_same_state(ow, w, "after %d frames" % upto)
    w.step(20)
    check_invariants(ow, w, n, dims)
    assert w.stats()["slow_path_steps"] == 0


def test_config3_pile_full_size():
    n, dims = 1 << 26, (10928, 8192)
    p = O.generate_scene(n, dims[0], dims[1], pile=True)
    ow, w = make_pair(dims, 3, p, capacity=max(n + 1024, 98495427))
    assert (ow.grid, ow.cells) == ((3643, 2731), 9949033)
    del p
    ow.step(1, threads=0)
    w.step(1)  # dense runs walked inside k_rebin
    assert_same_state(ow, w, "after 1 frame")
    ow.step(2, threads=0)
    w.step(2)  # ... and by k_rebin_dense
    assert_same_state(ow, w, "after 3 frames")
    check_invariants(ow, w, n, dims)
    assert w.stats()["slow_path_steps"] == 0


def test_config4_wide_world_full_size_one_gpu():
    """The 256 M-particle world of configs[4] on ONE device (the strips' single-GPU reference point)."""
    n, dims = 1 << 28, (65532, 5462)
    p = O.generate_scene(n, dims[0], dims[1])
    ow, w = make_pair(dims, 3, p)
    assert (ow.grid, ow.cells) == ((21845, 1821), 39779745)
    del p
    ow.step(2, threads=0)
    w.step(2)
    assert_same_state(ow, w, "after 2 frames")
    assert w.stats()["slow_path_steps"] == 0


# ---- BASELINE.json configs[1]: 1 M uniform, bit-exact after 1, 10, 100 frames ------------------

def test_config1_one_million_bit_exact():
    n, dims = 1 << 20, (1366, 1024)
    p = O.generate_scene(n, dims[0], dims[1])
    ow, w = make_pair(dims, 3, p)
    assert (ow.grid, ow.cells, ow.capacity) == ((456, 342), 155952, 1543928)
    done = 0
    for upto in (1, 10, 100):
        ow.step(upto - done, threads=0)
        w.step(upto - done)
        done = upto
        assert_same_state(ow, w, "after %d frames" % upto)
    assert w.stats()["slow_path_steps"] == 0


# ---- the host mirror driving the worker: WrachAPI / plugin systems ------------------------------

def test_wrach_api_flow_matches_oracle(kats):
    """runners/api/src/lib.rs:102-126 through the C++ WrachAPI, plus a seeded scene."""
    import wrach_b200 as W
    k = kats["api_smoke"]
    api = W.WrachAPI(W.WrachConfig(tuple(k["dimensions"]), cell_size=k["cell_size"]))
    api.add_particles(np.array(k["particles"], f32))
    for _ in range(k["ticks"]):
        api.tick()
    assert api.positions.shape[0] == k["readback_len"] == api.velocities.shape[0]
    assert tuple(api.positions[0]) != (0.0, 0.0) and tuple(api.velocities[0]) != (0.0, 0.0)

    dims, n = (400, 300), 70000
    p = O.generate_scene(n, dims[0], dims[1], seed=21)
    api = W.WrachAPI(W.WrachConfig(dims, cell_size=3))
    ow = O.OracleWorld(dims, 3)
    api.add_particles(p[: n // 2])
    ow.add_particles(p[: n // 2])
    for _ in range(3):
        api.tick()
        ow.step(1)
    # a second add_particles re-packs the STORE (not the GPU state) and re-uploads it: state.rs:90-101
    api.add_particles(p[n // 2:])
    ow.add_particles(p[n // 2:])
    for _ in range(3):
        api.tick()
        ow.step(1)
    ind, pos, vel = api.get_simulation_state().packed_data
    assert np.array_equal(ind, ow.indices)
    assert np.array_equal(pos, ow.positions_in) and np.array_equal(vel, ow.velocities_in)
    assert np.array_equal(api.positions, ow.positions_in)


# ---- seeded fuzz: many small worlds, degenerate grids included ---------------------------------

def _fuzz_cases():
    rng = np.random.default_rng(20261017)
    cases = []
    for i in range(40):
        cell = int(rng.choice([1, 2, 3, 3, 3, 4, 5, 7, 9]))
        gx = int(rng.choice([1, 2, 3, 5, 17, 64, 255, 256, 257, 300]))
        gy = int(rng.choice([1, 2, 3, 9, 40, 130]))
        w, h = gx * cell - int(rng.integers(0, cell)), gy * cell - int(rng.integers(0, cell))
        w, h = max(w, 1), max(h, 1)
        density = float(rng.choice([0.05, 0.3, 0.75, 0.75, 1.5, 4.0]))
        n = int(min(60000, max(1, density * w * h)))
        vscale = float(rng.choice([0.0, 0.5, 1.0, 1.0, 2.0, 30.0]))
        cases.append((i, (w, h), cell, n, vscale, int(rng.integers(0, 2))))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(), ids=lambda c: "fuzz%d-%dx%d-c%d-n%d-v%g" % (c[0], c[1][0], c[1][1], c[2], c[3], c[4]))
def test_fuzz_small_worlds(case):
    i, dims, cell, n, vscale, arith = case
    p = O.generate_scene(n, dims[0], dims[1], seed=1000 + i)
    p[:, 2:] *= f32(2.0 * vscale)  # |v| up to vscale
    ow, w = make_pair(dims, cell, p, arith=arith, capacity=2 * n + 64)
    for t in range(5):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "frame %d" % (t + 1))


# ---- SURVEY.md §8f "next" rows ------------------------------------------------------------------

def test_tick_active_reads_only_live_particles():
    import wrach_b200 as W
    dims, n = (200, 150), 20000
    p = O.generate_scene(n, dims[0], dims[1], seed=31)
    st = W.WrachState(W.WrachConfig(dims, cell_size=3))
    (gx, gy), total, cap = st.grid()
    st.add_particles(p)
    s0 = st.shader_settings.copy()
    s0.particles_in_frame_count = 0
    w = W.PhysicsComputeWorker(s0, total, cap)
    W.maybe_upload_to_gpu(w, st)
    ow = O.OracleWorld(dims, 3)
    ow.add_particles(p)
    w.step(4)
    ow.step(4)
    W.tick_active(w, st)
    ind, pos, vel = st.packed_data
    assert pos.shape[0] == n == vel.shape[0] and ind.shape[0] == total
    assert np.array_equal(ind, ow.indices)
    assert np.array_equal(pos, ow.positions_in[:n]) and np.array_equal(vel, ow.velocities_in[:n])
    W.tick(w, st)  # the reference's own semantics: capacity-sized
    assert st.packed_data[1].shape[0] == cap


def test_nonzero_view_anchor():
    """The shaders subtract view_anchor before keying (particles_per_cell.wgsl:14-15) and clamp to
    [anchor, anchor + dims] (particle.rs:46-52); the reference hard-wires (0,0) (builder.rs:61).
    With an anchor on a cell boundary the CPU packing and the device keys agree; check the step."""
    import ctypes
    import wrach_b200 as W
    from wrach_b200 import Buffers
    ax, ay, wdt, hgt, cell, n = 30.0, 60.0, 120.0, 90.0, 3, 9000
    rng = np.random.default_rng(5)
    p = np.empty((n, 4), f32)
    p[:, 0] = ax + rng.random(n, dtype=f32) * f32(wdt)
    p[:, 1] = ay + rng.random(n, dtype=f32) * f32(hgt)
    p[:, 2:] = rng.random((n, 2), dtype=f32) - f32(0.5)
    viewport = np.array([ax, ay, ax + wdt, ay + hgt], f32)
    _, (gx, gy) = O.active_grid(viewport, cell)
    total = gx * gy + 2
    ind = np.zeros(total, np.uint32)
    pos = np.zeros((n, 2), f32)
    vel = np.zeros((n, 2), f32)
    packed = O.lib().wo_create_packed_data(viewport, cell, p.reshape(-1), n, ind, pos.reshape(-1), vel.reshape(-1))
    assert packed == n
    s = O.Settings()
    s.view_dimensions[:] = [wdt, hgt]
    s.view_anchor[:] = [ax, ay]
    s.grid_dimensions[:] = [gx, gy]
    s.cell_size = cell
    s.particles_in_frame_count = n
    cap = 2 * n
    o_ind, o_pos, o_vel = ind.copy(), np.zeros((cap, 2), f32), np.zeros((cap, 2), f32)
    o_pos[:n], o_vel[:n] = pos, vel
    scratch_p, scratch_v = np.zeros((cap, 2), f32), np.zeros((cap, 2), f32)
    gs = W.WorldSettings()
    ctypes.memmove(ctypes.byref(gs), ctypes.byref(s), 32)
    g0 = gs.copy()
    g0.particles_in_frame_count = 0
    w = W.PhysicsComputeWorker(g0, total, cap)
    w.write_slice(Buffers.INDICES_MAIN, ind)
    w.write_slice(Buffers.POSITIONS_IN, pos)
    w.write_slice(Buffers.VELOCITIES_IN, vel)
    w.write(Buffers.WORLD_SETTINGS_UNIFORM, gs)
    for t in range(8):
        O.lib().wo_step(ctypes.byref(s), o_ind, o_pos.reshape(-1), o_vel.reshape(-1), scratch_p.reshape(-1),
                        scratch_v.reshape(-1), 1, O.ARITH_SPV)
        w.step(1)
        assert np.array_equal(w.read_vec(Buffers.INDICES_MAIN), o_ind), "frame %d" % (t + 1)
        assert np.array_equal(w.read_vec(Buffers.POSITIONS_IN)[:n], o_pos[:n])
        assert np.array_equal(w.read_vec(Buffers.VELOCITIES_IN)[:n], o_vel[:n])
    assert o_pos[:n, 0].min() >= ax and o_pos[:n, 0].max() <= ax + wdt


def test_viewport_streaming_window_onto_a_larger_world():
    """simulate a window, write it back into the store, move the window (anchor
    != 0), simulate on -- through the plugin calls, against the numpy + oracle restatement."""
    import wrach_b200 as W
    from tests.util import WindowedOracle
    rng = np.random.default_rng(11)
    n = 90000
    p = np.empty((n, 4), f32)
    p[:, 0] = rng.uniform(0, 900, n)
    p[:, 1] = rng.uniform(0, 300, n)
    p[:, 2:] = rng.uniform(-0.5, 0.5, (n, 2))
    state = W.WrachState(W.WrachConfig((300, 300), cell_size=3))
    state.add_particles(p)
    (gx, gy), total_cells, capacity = state.grid()
    s0 = state.shader_settings.copy()
    s0.particles_in_frame_count = 0
    worker = W.PhysicsComputeWorker(s0, total_cells, capacity)
    ref = WindowedOracle(p, (0, 0, 300, 300))
    for viewport in ((150, 0, 450, 300), (600, 0, 900, 300), (0, 0, 300, 300)):
        W.maybe_upload_to_gpu(worker, state)
        worker.step(4)
        W.tick_active(worker, state)
        ref.step(4)
        ind, pos, vel = state.packed_data
        assert np.array_equal(ind, ref.indices) and np.array_equal(pos, ref.pos[:ref.n]) and np.array_equal(vel, ref.vel[:ref.n])
        state.update_from_gpu()
        ref.update_from_gpu()
        state.set_viewport(viewport)
        ref.set_viewport(viewport)
    assert state.stored_particles == n
    W.maybe_upload_to_gpu(worker, state)
    worker.step(1)
    W.tick(worker, state)
    ref.step(1)
    ind, pos, vel = state.packed_data
    assert np.array_equal(ind, ref.indices) and np.array_equal(pos[:ref.n], ref.pos[:ref.n])


def test_push_division():
    """The pair push divides with a hand-written correctly rounded sequence (no range check, no
    slow-path call): bit-equal to div.rn for every divisor a push can see -- checked exhaustively."""
    import ctypes
    from wrach_b200 import _ffi
    bad = ctypes.c_ulonglong(123)
    _ffi.check(_ffi.lib().wrach_cuda_selftest_push_division(0, ctypes.byref(bad)))
    assert bad.value == 0


def test_push_square_root():
    """... and takes the distance with sqrt.rn's own fast path minus its range check and slow-path
    call (the tiny / zero case rides on one compare): bit-equal to sqrt.rn for every squared
    distance it is given -- checked exhaustively."""
    import ctypes
    from wrach_b200 import _ffi
    bad = ctypes.c_ulonglong(123)
    _ffi.check(_ffi.lib().wrach_cuda_selftest_push_sqrt(0, ctypes.byref(bad)))
    assert bad.value == 0


def test_plain_c_api_smoke_on_the_gpu(tmp_path):
    """examples/api_smoke.c -- the reference's API smoke test (runners/api/src/lib.rs:102-126) written
    against nothing but include/*.h -- built with gcc and run on the device: 3 coincident particles,
    5 ticks, capacity-sized read-back (164 positions)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, libdir = str(tmp_path / "api_smoke"), os.path.join(root, "wrach_b200", "lib")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "api_smoke.c"), "-L" + libdir, "-lwrach_cuda",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "read back 164 positions" in r.stdout


def _edge_scene(dims, cell, n, seed):
    """One particle per (random, distinct) cell, each placed a few floats either side of a cell edge
    and given the tiny velocity that lands it exactly on another float next to that edge: the move
    classification (exact compares in k_phys / k_tile_frame) against the reference's
    floor((x - anchor) / cell_size) (particles_per_cell.wgsl:14-27) where they could differ at all.
    Plus particles that leave the world through each side and come to rest exactly on its edge."""
    rng = np.random.default_rng(seed)
    gx, gy = int(dims[0] // cell) + 1, int(dims[1] // cell) + 1
    cells = rng.choice((gx - 2) * (gy - 2), size=n, replace=False)
    cx, cy = 1 + cells % (gx - 2), 1 + cells // (gx - 2)

    def near(edge, steps):
        v = edge.astype(f32)
        for s in range(3):
            v = np.where(steps > s, np.nextafter(v, f32(np.inf), dtype=f32), v)
            v = np.where(steps < -s, np.nextafter(v, f32(-np.inf), dtype=f32), v)
        return v.astype(f32)

    p = np.zeros((n, 4), f32)
    which = rng.integers(0, 3, n)          # 0: x edge, 1: y edge, 2: both (a corner)
    ex = (cx + rng.integers(0, 2, n)) * cell  # left or right edge of the cell
    ey = (cy + rng.integers(0, 2, n)) * cell
    x_from, x_to = near(ex, rng.integers(-3, 4, n)), near(ex, rng.integers(-3, 4, n))
    y_from, y_to = near(ey, rng.integers(-3, 4, n)), near(ey, rng.integers(-3, 4, n))
    mid_x = (cx * cell + rng.random(n) * cell).astype(f32)
    mid_y = (cy * cell + rng.random(n) * cell).astype(f32)
    on_x, on_y = which != 1, which != 0
    p[:, 0] = np.where(on_x, x_from, mid_x)
    p[:, 1] = np.where(on_y, y_from, mid_y)
    p[:, 2] = np.where(on_x, x_to - x_from, f32(0.01))   # exact: neighbouring floats (Sterbenz)
    p[:, 3] = np.where(on_y, y_to - y_from, f32(-0.01))
    leavers = np.array([[0.2, 10.0, -0.7, 0.0], [dims[0] - 0.2, 20.0, 0.9, 0.1], [30.0, 0.3, 0.0, -0.8],
                        [40.0, dims[1] - 0.1, 0.2, 0.6], [0.1, 0.1, -0.5, -0.5], [dims[0], dims[1], 0.3, 0.3]], f32)
    return np.concatenate([p, leavers])


@pytest.mark.parametrize("dims,cell", [((3000, 1800), 3), ((65532, 300), 3), ((4000, 900), 7), ((2000, 2000), 5), ((900, 700), 1)])
def test_cell_edges_one_float_either_side(dims, cell):
    p = _edge_scene(dims, cell, 20000, seed=dims[0] + cell)
    ow, w = make_pair(dims, cell, p, capacity=4 * len(p))
    for t in range(3):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))
