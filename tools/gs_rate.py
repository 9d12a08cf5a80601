import sys, os, numpy as np, torch
sys.path.insert(0,'.')
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
import tetsim_b200 as ts
from tetsim_b200 import mesh
m = mesh.load_dragon()
def rate(body, pp, frames=20):
    for _ in range(3): body.step(pp)
    body.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(frames): body.step(pp)
    e1.record(stream); torch.cuda.synchronize()
    return frames*pp["numSubsteps"]/e0.elapsed_time(e1)*1e3
p10=dict(ts.DEFAULT_PHYSICS_PARAMS,numSubsteps=10)
pw=dict(p10,worldBounds=list(mesh.wide_bounds(64.0)))
v64,t64=mesh.tile_bodies(m["tet_verts"],m["tet_ids"],8,8,y_shift=-0.40)
v784,t784=mesh.tile_bodies(m["tet_verts"],m["tet_ids"],28,28,y_shift=-0.40)
for th in sys.argv[1:]:
    os.environ["TETSIM_GS_THREADS"]=th
    b=ts.SoftBody(m["tet_verts"],m["tet_ids"],None,p10,solver="gs_exact",arithmetic="fast",stream=stream.cuda_stream)
    r1=rate(b,p10); b.close()
    b=ts.SoftBody(v64,t64,None,pw,solver="gs_exact",arithmetic="fast",stream=stream.cuda_stream); r64=rate(b,pw,10); b.close()
    b=ts.SoftBody(v784,t784,None,pw,solver="gs_exact",arithmetic="fast",stream=stream.cuda_stream); r784=rate(b,pw,10); b.close()
    print("threads %s: %.0f / %.0f / %.0f substeps/s (1 / 64 / 784 Dragons)" % (th, r1, r64, r784))
