"""Seeded benchmark scenes: the particle generator of examples/youre-a-pixel.rs:42-58 made
reproducible (the reference draws from an unseeded thread_rng).  Counter-based, so any slice of
particle ids can be generated independently (per strip, per chunk) and CPU and GPU legs see the same
bits:  u(id, c) = top 24 bits of splitmix64(seed ^ splitmix64(4*id + c)) * 2^-24
       x = u0*W   y = u1*H  (pile: y = H*u1^4)   vx = u2 - 0.5   vy = u3 - 0.5
BASELINE.md §3 fixes seed = 0x5752414348 and the world sizes of the five configurations."""
import numpy as np

SEED = 0x5752414348

# BASELINE.json configs -> (particles, world W x H, pile?)   cell_size = 3, anchor (0, 0)
WORKLOADS = {
    "1m-scene": dict(n=1_000_000, dims=(1480, 1052), pile=False),  # configs[0]
    "1m": dict(n=1 << 20, dims=(1366, 1024), pile=False),          # configs[1]
    "16m": dict(n=1 << 24, dims=(5464, 4096), pile=False),         # configs[2]  (HBM-roofline config)
    "64m-pile": dict(n=1 << 26, dims=(10928, 8192), pile=True),    # configs[3]
    "256m": dict(n=1 << 28, dims=(65532, 5462), pile=False),       # configs[4]
    "32m-strip": dict(n=1 << 25, dims=(8190, 5462), pile=False),   # one eighth of configs[4] (what one of 8 GPUs holds): tuning only
}


def _splitmix64(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _unit24(seed, ids, comp):
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) ^ _splitmix64(ids * np.uint64(4) + np.uint64(comp)))
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def generate(n, width, height, seed=SEED, first_id=0, pile=False, chunk=1 << 22):
    """(n, 4) float32 rows (x, y, vx, vy) for particle ids first_id .. first_id+n-1."""
    out = np.empty((n, 4), np.float32)
    w, h = np.float32(width), np.float32(height)
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        ids = np.arange(first_id + b, first_id + e, dtype=np.uint64)
        ux, uy = _unit24(seed, ids, 0), _unit24(seed, ids, 1)
        if pile:
            uy = (uy * uy) * (uy * uy)
        out[b:e, 0] = ux * w
        out[b:e, 1] = uy * h
        out[b:e, 2] = _unit24(seed, ids, 2) - np.float32(0.5)
        out[b:e, 3] = _unit24(seed, ids, 3) - np.float32(0.5)
    return out


def generate_fast(n, width, height, seed=SEED, first_id=0, pile=False, x0=0.0):
    """Same rows as generate() (checked in tests/test_host_mirror.py), produced by the C++ host
    library on all host threads; x is offset by x0 (strip workers generate their own columns)."""
    from . import api
    out = np.empty((n, 4), np.float32)
    api._lib().wrach_host_generate_scene(seed, first_id, n, x0, width, height, int(pile), out.ctypes.data)
    return out


def generate_columns(n, width, height, columns, cell=3, seed=SEED, pile=False, chunk=1 << 24):
    """The particles of the global scene (ids 0 .. n-1, as generate()) whose cell column lies in
    [columns[0], columns[1]), in ascending id order -- what a strip worker packs.  Every rank of a
    multi-GPU run calls this on the SAME scene, so the union over the strips is the whole scene
    whatever the number of strips (no per-strip generator, nothing lost on a strip edge).  The
    column is SpatialBin::get_cell_coord's (spatial_bin.rs:48-64): floor(x / cell) in f32."""
    c0, c1 = columns
    keep = []
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        p = generate_fast(m, width, height, seed=seed, first_id=b, pile=pile)
        cx = np.floor(p[:, 0] / np.float32(cell))
        keep.append(p[(cx >= c0) & (cx < c1)])
    return np.concatenate(keep) if keep else np.empty((0, 4), np.float32)


def state_checksum(indices, positions, velocities, columns, grid_x):
    """Order-sensitive 64-bit checksum of a packed strip (or of the whole world: columns = (0, gx)):
    sum over particles of a hash of (global cell, rank inside the cell, position bits, velocity
    bits), modulo 2^64 (wrach_host_packed_checksum).  The per-strip values add up to the same number
    however the world is cut, so bench lines at different GPU counts can be compared (same scene,
    same frame count)."""
    from . import api
    ind = np.ascontiguousarray(indices, np.uint32)
    pos = np.ascontiguousarray(positions, np.float32)
    vel = np.ascontiguousarray(velocities, np.float32)
    return int(api._lib().wrach_host_packed_checksum(ind.ctypes.data, ind.size, pos.ctypes.data, vel.ctypes.data,
                                                     columns[0], columns[1], grid_x))


def check_packed_invariants(indices, positions, velocities, columns, grid_x, dims, cell=3):
    """Size-independent properties of a packed frame read back from a worker (a strip or the whole
    world), checked by wrach_host_check_packed on all host threads: indices monotone with the two
    sentinels, every particle inside the slot range of the cell its position keys to (sortedness),
    positions inside the world, |v| <= 1.  Returns N; raises AssertionError naming the first
    property that fails."""
    from . import api
    ind = np.ascontiguousarray(indices, np.uint32)
    pos = np.ascontiguousarray(positions, np.float32)
    vel = np.ascontiguousarray(velocities, np.float32)
    rc = api._lib().wrach_host_check_packed(ind.ctypes.data, ind.size, pos.ctypes.data, vel.ctypes.data, columns[0],
                                            columns[1], grid_x, float(dims[0]), float(dims[1]), cell)
    if rc < 0:
        raise AssertionError({-1: "indices are not a monotone start table", -2: "a position lies outside the world",
                              -3: "|v| > 1 after a frame", -4: "a particle sits in a strip that does not own its column",
                              -5: "a particle is not in the slot range of the cell its position keys to"}.get(rc, "rc %d" % rc))
    return int(ind[-1])


def algorithmic_bytes(n, cells):
    """B = 64 N + 16 C per frame, split per kernel:
    physics: read (pos, vel) 16 + write 16 per particle, slot-range read 4 per cell;
    re-bin:  read 16 + write 16 per particle, counter write 4 + scan read/write 8 per cell."""
    return {"phys": 32 * n + 4 * cells, "rebin": 32 * n + 12 * cells, "step": 64 * n + 16 * cells}
