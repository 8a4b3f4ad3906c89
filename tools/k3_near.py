"""K3 count mode on near-ground-truth hypotheses only (ncu target): python tools/k3_near.py [near|random|mixed] [count|weighted]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
which = sys.argv[1] if len(sys.argv) > 1 else 'near'
mode = sys.argv[2] if len(sys.argv) > 2 else 'count'
prob = synth.make_problem(2000, 100000, 0.01, seed=1234)
T = synth.make_hypotheses(prob, 200000, seed=4321)
Ts = {'near': T[1::2], 'random': T[0::2], 'mixed': T[:100000]}[which]
e = PoseEngine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); e.set_stream(st.cuda_stream)
e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm)
Td = torch.from_numpy(np.ascontiguousarray(Ts).reshape(-1, 12)).cuda(); cd = torch.zeros(len(Ts), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(Ts), device='cuda')
for it in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
    print(which, mode, 'ms', a.elapsed_time(b))
