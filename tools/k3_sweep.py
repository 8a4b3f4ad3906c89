"""K3 kernel time over the CTA shapes (warps per CTA) on near-GT / random / mixed hypotheses.  python tools/k3_sweep.py"""
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
    for mode in ('count', 'weighted'):
        for nw, tab in ((32, -1), (32, 0), (32, 1), (24, -1)):
            e.set_option('k3_warps_' + mode, nw); e.set_option('k3_smem_table', tab)
            ms = []
            for it in range(6):
                flush.zero_()
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            got = cd.cpu().numpy()
            key = (name, mode)
            same = True if key not in ref else bool(np.array_equal(ref[key], got))
            ref.setdefault(key, got)
            print(f"{name:8s} {mode:8s} warps={nw} smem_table={tab} ms {min(ms[2:]):.4f} same_counts={same}", flush=True)
