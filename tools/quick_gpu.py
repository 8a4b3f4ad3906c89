import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
from oracle import pyoracle
prob = synth.make_problem(2000, 100000, 0.01)
T = synth.make_hypotheses(prob, 100000)
e = PoseEngine(0)
t=time.time(); e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm); print('setup', time.time()-t, e.grid_info())
o = pyoracle.PortOracle(prob.scene_xyz, prob.scene_nrm, prob.model_xyz, prob.model_nrm, prob.model_xyz, prob.model_nrm, prob.delta)
want = o.verify(T[:300])
c, s = e.score_lcp(0, T, 'count')
print('mismatch', int((c[:300] != want).sum()), c[:8], want[:8])
Td = torch.from_numpy(T.reshape(-1,12)).cuda(); cd = torch.zeros(len(T), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(T), device='cuda')
e.set_stream(torch.cuda.current_stream().cuda_stream)
for mode in ('count','weighted'):
    for it in range(3):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
        ms=a.elapsed_time(b); print(mode, 'ms', ms, 'hyp/s', len(T)/ms*1e3)
a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
a.record(); top = e.topk(0, 64); b.record(); torch.cuda.synchronize(); print('topk ms', a.elapsed_time(b), top['index'][:5], top['count'][:5])
