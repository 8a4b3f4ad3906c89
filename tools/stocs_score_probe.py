"""Why does scoring the StoCS hypotheses of configs[2] take longer than the operMode-0 ones?  python tools/stocs_score_probe.py"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
e = PoseEngine(0)
prob = synth.make_segment_problem(2000, 2000, 0.005, seed=100)
e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm); e.build_ppf_map(0)
print(e.grid_info())
for mode in (0, 1):
    n = e.generate_pcs(0, seed=5, max_hyp=200000, n_bases=2000, max_quads_per_base=100, mode=mode)
    for wmode in ('count', 'weighted', 'weighted'):
        torch.cuda.synchronize(); t0 = time.perf_counter(); e.score_generated(0, wmode); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('mode', mode, wmode, 'n', n, 'score ms', dt * 1e3, flush=True)
    t0 = time.perf_counter(); top = e.topk(0, 64); torch.cuda.synchronize(); print('  topk ms', (time.perf_counter() - t0) * 1e3)
    g = e.get_generated(0)
    T = g[0] if isinstance(g, tuple) else g['T']
    T = np.asarray(T).reshape(-1, 3, 4)
    mv = prob.model_xyz - prob.model_xyz.mean(0)
    rinf = np.abs(mv).max()
    bound = (np.abs(T[:, :, :3]).sum(2) * rinf + np.abs(T[:, :, 3])).max(1)
    print('  |t| median', np.median(np.linalg.norm(T[:, :, 3], axis=1)), 'bound median', np.median(bound), 'frac bound > 0.3:', (bound > 0.3).mean(), 'top score', top['score'][:3])
print(e.label_stats())
