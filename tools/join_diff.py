"""Dumps the quads on which the device join and the reference's FindCongruentQuadrilaterals disagree (tests/golden/pcs_small.npz)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from physimglobalpose_b200.engine import PoseEngine
g = np.load('tests/golden/pcs_small.npz')
e = PoseEngine(0)
delta = float(g['delta'])
e.set_scene(g['scene_xyz'], g['scene_nrm'], delta); e.set_model(0, g['model_xyz'], g['model_nrm'])
out = {}
tot = [0, 0, 0]
for k, (b, inv) in enumerate(zip(g['bases'], g['invariants'])):
    q = e.find_quads(0, b, inv[0], inv[1], delta, g[f'b{k}_p1'], g[f'b{k}_p2'])
    ours = set(map(tuple, q.tolist())); ref = set(map(tuple, g[f'b{k}_quads'].tolist()))
    out[f'b{k}_only_ours'] = np.array(sorted(ours - ref), np.int32).reshape(-1, 4)
    out[f'b{k}_only_ref'] = np.array(sorted(ref - ours), np.int32).reshape(-1, 4)
    tot[0] += len(ref); tot[1] += len(ours - ref); tot[2] += len(ref - ours)
    print(k, len(ref), len(ours), len(ours - ref), len(ref - ours))
print('total ref', tot[0], 'only ours', tot[1], 'only ref', tot[2])
np.savez('gpurun_out/join_diff.npz', **out)
