import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
from oracle import pyoracle
nm, ns, nh, delta = [int(x) for x in sys.argv[1:4]] + [float(sys.argv[4])] if len(sys.argv) > 4 else (2000, 100000, 100000, 0.01)
prob = synth.make_problem(nm, ns, delta)
T = synth.make_hypotheses(prob, nh)
e = PoseEngine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); e.set_stream(st.cuda_stream)
t=time.time(); e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm); torch.cuda.synchronize(); print('setup', time.time()-t, e.grid_info())
t=time.time(); e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); torch.cuda.synchronize(); print('set_scene again', time.time()-t)
o = pyoracle.PortOracle(prob.scene_xyz, prob.scene_nrm, prob.model_xyz, prob.model_nrm, prob.model_xyz, prob.model_nrm, prob.delta)
nchk = min(nh, 400)
want = o.verify(T[:nchk])
c, s = e.score_lcp(0, T, 'count')
print('mismatch', int((c[:nchk] != want).sum()), c[:8], want[:8])
Td = torch.from_numpy(T.reshape(-1,12)).cuda(); cd = torch.zeros(len(T), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(T), device='cuda')
for mode, fc in (('count',0),('count',1),('weighted',0),('weighted',1)):
    e.set_option('force_coarse', fc)
    for it in range(3):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
        ms=a.elapsed_time(b)
    print(mode, 'force_coarse', fc, 'ms', ms, 'hyp/s', len(T)/ms*1e3)
    if mode == 'count':
        print('  equal to host-api result', bool((cd.cpu().numpy().astype(np.uint32) == c).all()))
a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
a.record(); top = e.topk(0, 64); b.record(); torch.cuda.synchronize(); print('topk ms', a.elapsed_time(b), top['index'][:5], top['count'][:5])
