import sys
sys.path.insert(0,'.')
import tetsim_b200 as ts
from tetsim_b200 import mesh
m = mesh.load_dragon()
p10=dict(ts.DEFAULT_PHYSICS_PARAMS,numSubsteps=10)
b=ts.SoftBody(m["tet_verts"],m["tet_ids"],None,p10,solver="gs_exact",arithmetic="fast")
for _ in range(6): b.simulate(1/600, p10)
b.synchronize()
