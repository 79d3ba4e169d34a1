"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(tests/golden/reference_kats.json, each entry cites the reference file:line) and against the
SPIR-V arithmetic fixture (tests/golden/spv_arith.json, produced by oracle/tools/spv_dis.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


@pytest.mark.parametrize("arith", [O.ARITH_UNFUSED, O.ARITH_SPV])
def test_pair_push_golden_vector(kats, arith):
    k = kats["pair_push"]
    got = O.pairs(k["positions"], arith)
    # Vec2 equality in the reference is exact f32 equality against these decimal literals
    assert np.array_equal(got, np.array(k["expected_positions"], f32))
    d = f32(np.sqrt(np.sum((got[0].astype(np.float64) - got[1]) ** 2)))
    assert k["new_distance_between"][0] < d < k["new_distance_between"][1]


def test_cell_coord(kats):
    for k in kats["cell_coord"]:
        got = [O.cell_coord(k["position"][0], k["cell_size"]), O.cell_coord(k["position"][1], k["cell_size"])]
        assert got == k["coord"], k["cite"]


def test_active_cells(kats):
    for k in kats["active_cells"]:
        cells, grid = O.active_cells(k["viewport"], k["cell_size"])
        assert [list(c) for c in cells] == k["cells"], k["cite"]
        if k["grid"] is not None:
            assert list(grid) == k["grid"], k["cite"]


def _world_for_viewport(viewport, cell_size):
    assert viewport[0] == 0.0 and viewport[1] == 0.0
    return O.OracleWorld((viewport[2], viewport[3]), cell_size)


def test_packed_data(kats):
    for k in kats["packed_data"]:
        w = _world_for_viewport(k["viewport"], k["cell_size"])
        indices, pos, vel = w.pack(np.array(k["particles"], f32))
        assert indices.tolist() == k["indices"], k["cite"]
        assert np.array_equal(pos, np.array(k["positions"], f32)), k["cite"]
        assert np.array_equal(vel, np.array(k["velocities"], f32)), k["cite"]


def test_capacity(kats):
    for k in kats["capacity"]:
        w = _world_for_viewport(k["viewport"], k["cell_size"])
        assert w.capacity == k["max_particles_per_frame"], k["cite"]


@pytest.mark.parametrize("threads", [1, 3])
def test_indices_after_ticks_equal_cpu_packing(kats, threads):
    """03_prefix_sum.rs:151-260: after 4 ticks of static particles the device `indices` equal
    create_packed_data().indices."""
    for k in kats["gpu_equals_cpu_indices"]:
        w = O.OracleWorld(k["dimensions"], k["cell_size"])
        p = np.array(k["particles"], f32)
        w.add_particles(p)
        cpu_indices, _, _ = w.pack(p)
        if k["indices"] is not None:
            assert cpu_indices.tolist() == k["indices"], k["cite"]
        else:
            assert w.total_cells == k["total_cells"], k["cite"]
        for _ in range(k["ticks"]):
            w.step(1, threads=threads)
        assert np.array_equal(w.indices, cpu_indices), k["cite"]


def test_packed_positions_after_ticks(kats):
    k = kats["packed_positions_after_ticks"]
    w = O.OracleWorld(k["dimensions"], k["cell_size"])
    p = np.array(k["particles"], f32)
    w.add_particles(p)
    _, cpu_pos, _ = w.pack(p)
    assert np.array_equal(cpu_pos, np.array(k["positions"], f32))
    for _ in range(k["ticks"]):
        w.step()
    a, b = k["first_cell_slots"]
    got = w.positions_in[:4].copy()
    first = got[a:b][np.argsort(got[a:b, 0])]  # in-cell order is free in the reference's assertion
    assert np.array_equal(first, np.array(k["positions"][a:b], f32))
    assert np.array_equal(got[2:4], np.array(k["positions"][2:4], f32))
    # our canonical order is the stable one, so here even the unsorted slice matches
    assert np.array_equal(got, np.array(k["positions"], f32))


@pytest.mark.parametrize("arith", [O.ARITH_UNFUSED, O.ARITH_SPV])
def test_api_smoke(kats, arith):
    k = kats["api_smoke"]
    w = O.OracleWorld(k["dimensions"], k["cell_size"], arith=arith)
    w.add_particles(np.array(k["particles"], f32))
    for _ in range(k["ticks"]):
        w.step()
    assert w.positions_in.shape[0] == k["readback_len"] == w.velocities_in.shape[0]
    assert tuple(w.positions_in[0]) != (0.0, 0.0)
    assert tuple(w.velocities_in[0]) != (0.0, 0.0)
    # three coincident particles: distance == 0 -> 0.0001, delta == 0, so they just translate
    assert np.array_equal(w.positions_in[:3], np.full((3, 2), 7.5, f32))


def test_spv_arith_fixture_matches_oracle_spv_mode():
    """The shipped SPIR-V fuses exactly: dist^2 = fma(dx,dx,dy*dy) with d = left-right, and the four
    position updates fma(-/+(right-left), force, pos); force = (0.5*(1-dist))/dist stays unfused."""
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "spv_arith.json")))
    assert fx["glsl_ext_histogram"] == {"Fma": 5, "Sqrt": 1}
    assert fx["float_op_histogram"]["FDiv"] == 1
    assert fx["local_size"] == [32, 1, 1]
    fm = fx["fma_expressions"]
    assert fm[0].startswith("fma((") and ".x - " in fm[0] and fm[0].count(".y") == 4
    assert fm[1].startswith("fma(-(") and fm[2].startswith("fma(-(")
    assert fm[3].startswith("fma((") and fm[4].startswith("fma((")
    assert all("((0.5 * (1.0 - " in e and ") / " in e for e in fm[1:])
    assert fx["sqrt_expressions"][0] == "sqrt(%s)" % fm[0]
    # the two modes differ somewhere (otherwise the distinction would be untestable) ...
    rng = np.random.default_rng(7)
    differs = False
    for _ in range(200):
        p = (rng.random((9, 2), dtype=f32) * f32(3.0)).astype(f32)
        a, b = O.pairs(p, O.ARITH_UNFUSED), O.pairs(p, O.ARITH_SPV)
        differs |= not np.array_equal(a, b)
        assert np.allclose(a, b, atol=1e-5)  # ... but only by rounding
    assert differs
