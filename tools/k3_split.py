"""Where K3's time goes: near-ground-truth hypotheses vs random ones, cull on / off.  python tools/k3_split.py"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
prob = synth.make_problem(2000, 100000, 0.01, seed=1234)
T = synth.make_hypotheses(prob, 200000, seed=4321)
e = PoseEngine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); e.set_stream(st.cuda_stream)
e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
sets = {'mixed': T[:100000], 'near-GT': T[1::2], 'random': T[0::2]}
for name, Ts in sets.items():
    Td = torch.from_numpy(np.ascontiguousarray(Ts).reshape(-1, 12)).cuda(); cd = torch.zeros(len(Ts), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(Ts), device='cuda')
    for mode in ('count', 'weighted'):
        for cull in (1, 0):
            e.set_option('group_cull', cull)
            ms = []
            for it in range(6):
                flush.zero_()
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            print(f"{name:8s} {mode:8s} cull={cull} ms {min(ms[2:]):.4f} mean count {cd.float().mean().item():.1f}", flush=True)
print(e.label_stats())
