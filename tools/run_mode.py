"""Runs K3 in one mode a few times on the C2 inputs (for ncu captures): python tools/run_mode.py weighted|count [n_model n_scene n_hyp delta]"""
import sys
sys.path.insert(0, '.')
import torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
mode = sys.argv[1] if len(sys.argv) > 1 else "weighted"
nm, ns, nh = [int(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else (2000, 100000, 100000)
delta = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
prob = synth.make_problem(nm, ns, delta, seed=1234)
T = synth.make_hypotheses(prob, nh, seed=4321)
e = PoseEngine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); e.set_stream(st.cuda_stream)
e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm)
Td = torch.from_numpy(T.reshape(-1, 12)).cuda(); cd = torch.zeros(len(T), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(T), device='cuda')
for it in range(4):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
    print(mode, 'ms', a.elapsed_time(b), 'hyp/s', len(T) / a.elapsed_time(b) * 1e3)
print(e.grid_info())
print(e.label_stats())
