#!/usr/bin/env python3
"""Same-box A/B timing of library variants and runtime knobs on one prepared scene (GPU box).

  python tools/ab.py 16m 200 3  base=wrach_b200/lib/sweep/lib_base.so  new=  pdl=:WRACH_PDL=1

Every argument after (workload, frames, repeats) is  name=[library path][:ENV=VALUE[,ENV=VALUE...]];
an empty path means the product library.  The scene is generated and packed ONCE (host mirror),
saved under /dev/shm, and each variant is timed in its own process (the library is chosen at import
time), the variants interleaved repeat by repeat so that clock drift hits them alike.  Prints one
line per (repeat, variant): frame time by CUDA events over the batch, and the k_phys / re-bin split.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def prepare(workload):
    import wrach_b200 as W
    from wrach_b200 import scene
    path = "/dev/shm/wrach_ab_%s.npz" % workload
    if os.path.exists(path):
        return path
    wl = scene.WORKLOADS[workload]
    state = W.WrachState(W.WrachConfig(wl["dims"], cell_size=3))
    state.add_particles(scene.generate_fast(wl["n"], wl["dims"][0], wl["dims"][1], pile=wl["pile"]))
    (gx, gy), total_cells, capacity = state.grid()
    ind, pos, vel = state.create_packed_data()
    s = state.shader_settings
    np.savez(path, ind=ind, pos=pos, vel=vel, total_cells=total_cells, capacity=max(capacity, wl["n"]),
             settings=np.frombuffer(bytes(s), np.uint8))
    return path


def child(path, frames):
    import ctypes
    import wrach_b200 as W
    from wrach_b200 import Buffers
    d = np.load(path)
    s = W.WorldSettings.from_buffer_copy(d["settings"].tobytes())
    create = s.copy()
    create.particles_in_frame_count = 0
    w = W.PhysicsComputeWorker(create, int(d["total_cells"]), int(d["capacity"]))
    w.write_slice(Buffers.INDICES_MAIN, d["ind"])
    w.write_slice(Buffers.POSITIONS_IN, d["pos"])
    w.write_slice(Buffers.VELOCITIES_IN, d["vel"])
    w.write(Buffers.WORLD_SETTINGS_UNIFORM, s)
    w.step_timed(10)
    ms = w.step_timed(frames) / frames
    pf = min(frames, 30)
    a, b = w.step_profiled(pf)
    ind = w.read_vec(Buffers.INDICES_MAIN)
    pos = w.read_vec(Buffers.POSITIONS_IN)[:int(ind[-1])].view(np.uint32).astype(np.uint64)
    pcrc = int(np.bitwise_xor.reduce((pos[:, 0] * np.uint64(0x9E3779B1) + pos[:, 1]) * np.arange(1, pos.shape[0] + 1, dtype=np.uint64)))
    st = w.stats()
    print(json.dumps({"ms": ms, "phys": a / pf, "rebin": b / pf, "n": int(ind[-1]), "slow": st["slow_path_steps"], "tiles": st["tile_frames"], "fb": st["tile_fallbacks"],
                      "crc": int(np.bitwise_xor.reduce(ind.astype(np.uint64) * np.arange(1, ind.size + 1, dtype=np.uint64))) ^ pcrc}))
    w.close()


def main():
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        return
    workload, frames, repeats = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    path = prepare(workload)
    variants = []
    for spec in sys.argv[4:]:
        name, rest = spec.split("=", 1)
        lib, _, envs = rest.partition(":")
        env = dict(os.environ)
        if lib:
            env["WRACH_CUDA_LIB"] = os.path.join(ROOT, lib)
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("=")
            env[k] = v
        variants.append((name, env))
    for rep in range(repeats):
        for name, env in variants:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path, str(frames)], env=env,
                               capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            try:
                d = json.loads(line)
                print("%-10s rep %d  frame %.4f ms  phys %.4f  rebin+scan %.4f  n %d slow %d tile frames %d fallbacks %d crc %x" % (
                    name, rep, d["ms"], d["phys"], d["rebin"], d["n"], d["slow"], d["tiles"], d["fb"], d["crc"]), flush=True)
            except Exception:
                print("%-10s rep %d  FAILED rc=%d %s" % (name, rep, r.returncode, (r.stderr or line)[-300:]), flush=True)


if __name__ == "__main__":
    main()
