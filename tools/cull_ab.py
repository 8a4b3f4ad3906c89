"""A/B of K3's group cull on the bench inputs: python tools/cull_ab.py [n_model n_scene n_hyp delta]
Prints kernel time with the cull on / off (count and weighted), and checks the counts against the C oracle on a prefix."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
from oracle import pyoracle

nm, ns, nh = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (2000, 100000, 100000)
delta = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
prob = synth.make_problem(nm, ns, delta, seed=1234)
T = synth.make_hypotheses(prob, nh, seed=4321)
e = PoseEngine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); e.set_stream(st.cuda_stream)
e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm)
Td = torch.from_numpy(T.reshape(-1, 12)).cuda(); cd = torch.zeros(len(T), dtype=torch.int32, device='cuda'); sd = torch.zeros(len(T), device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
pyoracle.build_port()
o = pyoracle.PortOracle(prob.scene_xyz, prob.scene_nrm, prob.model_xyz, prob.model_nrm, prob.model_xyz, prob.model_nrm, prob.delta)
nchk = min(nh, 1500 if nm <= 4000 else 100)
want = o.verify(T[:nchk])
ws, wn = o.weighted_verify(T[:nchk])
res = {}
for mode in ('count', 'weighted'):
    for cull in (1, 0):
        e.set_option('group_cull', cull)
        ms = []
        for it in range(6):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); e.score_lcp_device(0, Td, cd, sd, mode); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        got = cd.cpu().numpy().astype(np.uint32)
        res[(mode, cull)] = got
        ref = want if mode == 'count' else wn.astype(np.uint32)
        print(f"{mode:8s} cull={cull} ms {min(ms[2:]):.4f} (median {sorted(ms[2:])[2]:.4f})  hyp/s {len(T) / min(ms[2:]) * 1e3:.4g}  mismatches vs oracle {int((got[:nchk] != ref).sum())}/{nchk}", flush=True)
    print(f"{mode:8s} cull on == off over all {len(T)}: {bool(np.array_equal(res[(mode, 1)], res[(mode, 0)]))}", flush=True)
print(e.grid_info())
