"""Debug tool: per-batch frame time of the 64M pile (the scene diffuses, so the time drops frame by frame)."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wrach_b200 as W
from wrach_b200 import scene
wl=scene.WORKLOADS['64m-pile']
state=W.WrachState(W.WrachConfig(wl['dims'],cell_size=3))
(gx,gy),tc,cap=state.grid()
state.add_particles(scene.generate_fast(wl['n'],*wl['dims'],pile=True))
s0=state.shader_settings.copy(); s0.particles_in_frame_count=0
w=W.PhysicsComputeWorker(s0,tc,max(cap,wl['n']))
W.maybe_upload_to_gpu(w,state); w.sync()
for i in range(6):
    ms=w.step_timed(5); print('batch',i,'ms/step %.3f'%(ms/5), w.stats())
for i in range(3):
    a,b=w.step_profiled(5); print('prof',i,a/5,b/5)
