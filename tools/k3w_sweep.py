"""Weighted K3 kernel time over the tuning hooks (warps per CTA, bmrank table in shared memory or through L1).  python tools/k3w_sweep.py"""
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
ref = {}
for name, Ts in sets.items():
    Td = torch.from_numpy(np.ascontiguousarray(Ts).reshape(-1, 12)).cuda(); cd = torch.zeros(len(Ts), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(Ts), device='cuda')
    for nw in (32, 28, 24):
        for tab in (0, 1):
            e.set_option('k3_warps_weighted', nw); e.set_option('k3_smem_table', tab)
            ms = []
            for it in range(5):
                flush.zero_()
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); e.score_lcp_device(0, Td, cd, sd, 'weighted'); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            got = sd.cpu().numpy()
            same = True if name not in ref else bool(np.array_equal(ref[name], got))
            ref.setdefault(name, got)
            print(f"{name:8s} warps={nw} smem_table={tab} ms {min(ms[2:]):.4f} same={same}", flush=True)
