#!/usr/bin/env python3
"""Warp-instruction model of k_phys's pair phase on REAL push patterns (CPU only).

The pair phase is 34 % of k_phys's instructions at 14.6 of 32 lanes: a push fires for
~22 % of pairs, so the 29-instruction push path runs in almost every slot for a quarter of the
lanes.  This replays a steady-state frame of a uniform scene (the checker steps it, pair_trace.c
records which pairs push in every cell) under alternative schedules and counts warp instructions,
calibrated on the measured kernel:

  current    thread per cell, cells of a run sorted by occupancy, unrolled partner loop:
             a slot costs T (test) for the warp, plus P when any lane pushes
  scan-K     each lane walks its own pair list: up to K tests, then one shared push round
             (a lane stops at its first pushing pair); test costs T2 (dynamic pair index)
  two-cell   a lane interleaves two cells: two tests per slot and ONE shared push per slot
             (on a conflict the second cell waits a slot)

Usage: python tests/tools/model/pair_phase_model.py [particles=1048576] [frames=20]
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (lives under tests/: only test infrastructure may use the checker)

T, P = 11, 30      # current kernel: test + loop control, push path (SASS counts)
T2 = 15            # a test whose pair index is dynamic (table / incremental advance)


def trace(n, frames):
    so = os.path.join(HERE, "libpairtrace.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "pair_trace.c"), "-lm"])
    lib = ctypes.CDLL(so)
    side = int((n / 0.75) ** 0.5)
    dims = (side * 4 // 3, side * 3 // 4)
    w = O.OracleWorld(dims, 3)
    w.add_particles(O.generate_scene(n, dims[0], dims[1], seed=7))
    w.step(frames, threads=0)
    n9 = np.zeros(w.cells, np.uint8)
    mask = np.zeros(w.cells, np.uint64)
    lib.pair_trace(w.indices.ctypes.data_as(ctypes.c_void_p), w.positions_in.ctypes.data_as(ctypes.c_void_p),
                   ctypes.c_uint32(w.cells), n9.ctypes.data_as(ctypes.c_void_p), mask.ctypes.data_as(ctypes.c_void_p))
    return w, n9, mask


def pair_bits(n9, mask):
    """list of per-pair push flags in reference order for one cell"""
    k = n9 * (n9 - 1) // 2
    return [(int(mask) >> b) & 1 for b in range(k)]


def model_current(n9, mask):
    """runs of 256 cells, sorted by n9 descending, warp = 32 consecutive sorted cells"""
    total = tests = pushes_run = 0
    cells = len(n9)
    for r0 in range(0, cells, 256):
        idx = np.arange(r0, min(r0 + 256, cells))
        order = idx[np.argsort(-n9[idx].astype(np.int32), kind="stable")]
        for w0 in range(0, len(order), 32):
            lanes = order[w0:w0 + 32]
            nmax = int(n9[lanes].max())
            if nmax < 2:
                continue
            # slot (i, u) exists for the warp if any lane has n9 > i + u ... iterate rows / partners
            for i in range(nmax - 1):
                for u in range(1, nmax - i):
                    any_push = False
                    for c in lanes:
                        n = int(n9[c])
                        if i + u < n:
                            # bit index of pair (i, i+u) in a cell of n particles
                            b = i * n - i * (i + 1) // 2 + (u - 1)
                            if (int(mask[c]) >> b) & 1:
                                any_push = True
                                break
                    tests += 1
                    pushes_run += any_push
    total = tests * T + pushes_run * P
    return total, tests, pushes_run


def model_scan(n9, mask, K):
    total = 0
    cells = len(n9)
    for r0 in range(0, cells, 256):
        idx = np.arange(r0, min(r0 + 256, cells))
        order = idx[np.argsort(-n9[idx].astype(np.int32), kind="stable")]
        for w0 in range(0, len(order), 32):
            lanes = order[w0:w0 + 32]
            seqs = [pair_bits(int(n9[c]), mask[c]) for c in lanes]
            pos = [0] * len(seqs)
            while any(pos[q] < len(seqs[q]) for q in range(len(seqs))):
                steps = 0
                found_any = False
                for q, s in enumerate(seqs):
                    t = 0
                    while pos[q] < len(s) and t < K:
                        t += 1
                        pos[q] += 1
                        if s[pos[q] - 1]:
                            found_any = True
                            break
                    steps = max(steps, t)
                total += steps * T2 + (P + 2 if found_any else 0)
    return total


def model_two_cell(n9, mask):
    """lane = two cells (sorted order: cell q and cell q+128 of the run -> 4 warps per run)"""
    total = 0
    cells = len(n9)
    for r0 in range(0, cells, 256):
        idx = np.arange(r0, min(r0 + 256, cells))
        order = idx[np.argsort(-n9[idx].astype(np.int32), kind="stable")]
        half = (len(order) + 1) // 2
        a_cells, b_cells = order[:half], order[half:][::-1]  # heavy with light
        for w0 in range(0, half, 32):
            la = a_cells[w0:w0 + 32]
            lb = b_cells[w0:w0 + 32]
            sa = [pair_bits(int(n9[c]), mask[c]) for c in la]
            sb = [pair_bits(int(n9[c]), mask[c]) for c in lb] + [[]] * (len(sa) - len(lb))
            pa, pb = [0] * len(sa), [0] * len(sa)
            while any(pa[q] < len(sa[q]) or pb[q] < len(sb[q]) for q in range(len(sa))):
                any_push = False
                for q in range(len(sa)):
                    a_has, b_has = pa[q] < len(sa[q]), pb[q] < len(sb[q])
                    a_push = a_has and sa[q][pa[q]]
                    b_push = b_has and sb[q][pb[q]]
                    if a_has:
                        pa[q] += 1
                    if b_has and not (a_push and b_push):
                        pb[q] += 1
                    any_push |= bool(a_push or b_push)
                total += 2 * T2 + 8 + (P + 4 if any_push else 0)
    return total


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    w, n9, mask = trace(n, frames)
    pairs = int((n9.astype(np.int64) * (n9.astype(np.int64) - 1) // 2).sum())
    pushes = int(sum(bin(int(m)).count("1") for m in mask))
    print("%d particles, %d cells, %.2f particles/cell, %d pairs/frame, push rate %.3f" % (
        w.n, w.cells, w.n / w.cells, pairs, pushes / max(pairs, 1)))
    cur, tests, pr = model_current(n9, mask)
    per32 = 32.0 / w.n
    print("current : %8.1f warp-instr per 32 particles  (%d slots, push path taken in %.1f %% of them)" % (
        cur * per32, tests, 100.0 * pr / max(tests, 1)))
    print("          measured: 72.1 M for 16.7 M particles = 137.5 per 32 particles")
    for K in (2, 3, 4, 6):
        s = model_scan(n9, mask, K)
        print("scan-%d  : %8.1f  (%+.0f %%)" % (K, s * per32, 100.0 * (s - cur) / cur))
    t = model_two_cell(n9, mask)
    print("two-cell: %8.1f  (%+.0f %%)" % (t * per32, 100.0 * (t - cur) / cur))
    ideal = (pairs * 8 + pushes * 28) / 32.0
    print("every lane busy (lower bound): %.1f" % (ideal * per32))


if __name__ == "__main__":
    main()
