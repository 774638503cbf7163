import numpy as np, torch, sys
sys.path.insert(0,'.')
from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
from oracle.c_oracle import COracle
from oracle import edgecheck_oracle as O
np.set_printoptions(precision=5, suppress=True, linewidth=200)
model=load_iiwa14(); worlds=synthetic_worlds(64)
chk=EdgeChecker(model,worlds,EdgeCheckConfig(max_steps=32))
q0,q1=synthetic_edges(model,1<<17)
e=16*32
Q=O.edge_configs(model,q0[e].astype(np.float64),q1[e].astype(np.float64),np.radians(3))
q=Q[5:6].astype(np.float32)
g=chk.check_configs(torch.from_numpy(q))
c=COracle(model,worlds).check_configs(q)
w=31
print('gpu defl',g.deflection.cpu().numpy()[0,w]); print('c   defl',c['deflection'][0,w])
print('gpu theta',g.theta.cpu().numpy()[0,w].ravel()); print('c   theta',c['theta'][0,w].ravel())
r=O.check_configs(model,[worlds[w]],q.astype(np.float64),O.OracleParams())
print('np  defl',r.deflection[0,0]); print('np theta',r.theta[0,0].ravel())
# edge kernel masks for that edge
out=chk.validate_edges(torch.from_numpy(q0[e:e+1]),torch.from_numpy(q1[e:e+1]),want_masks=True)
ce=COracle(model,worlds).validate_edges(q0[e:e+1],q1[e:e+1],max_steps=32,want_masks=True)
print('edge exceed gpu %08x c %08x'%(int(out.exceed_mask.cpu().numpy().view(np.uint32)[0,w,0]), int(ce['exceed_mask'][0,w,0])))
print('maxdefl',float(out.max_deflection[0,w]), ce['max_deflection'][0,w])
# all steps deflection from config kernel
gq=chk.check_configs(torch.from_numpy(Q.astype(np.float32))); cq=COracle(model,worlds).check_configs(Q.astype(np.float32))
print('per-step max defl gpu',gq.deflection.cpu().numpy()[:,w].max(1)); print('per-step max defl c  ',cq['deflection'][:,w].max(1))
