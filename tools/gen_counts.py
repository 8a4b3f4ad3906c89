import sys, time
sys.path.insert(0, '.')
import torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
prob = synth.make_segment_problem(2000, 2000, 0.005, seed=100)
e = PoseEngine(0)
e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta); e.set_model(0, prob.model_xyz, prob.model_nrm)
for mode in (0,):
    for mq in (100, 0):
        e.generate_pcs(0, seed=1, max_hyp=1000, n_bases=8, mode=mode)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = e.generate_pcs(0, seed=5, max_hyp=50_000_000, n_bases=100, max_quads_per_base=mq, mode=mode)
        torch.cuda.synchronize(); print('mode', mode, 'max_quads', mq, 'hyps', n, 'ms', (time.perf_counter() - t0) * 1e3)
ids, inv, ok = e.get_bases(0)
import numpy as np
P = prob.scene_xyz - e.centroids(0)[0]
for b in range(5):
    d1 = float(np.linalg.norm(P[ids[b,0]] - P[ids[b,1]])); d2 = float(np.linalg.norm(P[ids[b,2]] - P[ids[b,3]]))
    p1 = e.extract_pairs(0, d1, prob.delta); p2 = e.extract_pairs(0, d2, prob.delta)
    q = e.find_quads(0, ids[b], inv[b,0], inv[b,1], prob.delta, p1, p2, cap=1 << 24)
    print('base', b, 'd', round(d1,3), round(d2,3), 'pairs', len(p1), len(p2), 'quads', len(q))
