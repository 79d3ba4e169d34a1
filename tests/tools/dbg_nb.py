import sys, numpy as np
sys.path.insert(0, "/root/repo")
import wrach_b200 as W
from oracle import oracle as O
from tests.test_gpu_strips import make_strips
dims, n = (240, 150), 40000
p = O.generate_scene(n, dims[0], dims[1], seed=55)
for ns in (1, 2):
    ow = O.OracleWorld(dims, 3, neighbours=True); ow.add_particles(p); ow.step(1)
    workers, columns, grid = make_strips(dims, 3, p, ns)
    for w in workers: w.set_neighbour_mode(True)
    W.PhysicsComputeWorker.strip_group_step(workers, 1)
    gx, gy = grid
    oind = ow.indices.astype(np.int64)
    for w,(c0,c1) in zip(workers, columns):
        ind = w.read_vec(W.Buffers.INDICES_MAIN).astype(np.int64)
        width = c1-c0
        lcnt = (ind[2:]-ind[1:-1]).reshape(gy,width)
        gcell = (np.arange(gy)[:, None] * gx + np.arange(c0, c1)[None, :])
        gcnt = (oind[2:][gcell]-oind[1:][gcell])
        bad = np.argwhere(lcnt!=gcnt)
        print("strips", ns, "strip", c0,c1, "bad cells", len(bad), "rows hist", np.bincount(bad[:,0], minlength=gy) if len(bad) else None, "sum diff", (lcnt-gcnt).sum(), w.stats()["slow_path_steps"])
        if len(bad):
            for b in bad[:6]: print("   cell row %d col %d: got %d oracle %d" % (b[0], b[1], lcnt[b[0],b[1]], gcnt[b[0],b[1]]))
    for w in workers: w.close()
