"""TrICP of the top-64 generated poses on a 2k-pt segment / 2k-pt model request (ncu target + timing): python tools/tricp_run.py"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine
seg = synth.make_segment_problem(2000, 2000, 0.005, seed=5)
e = PoseEngine(0)
e.set_scene(seg.scene_xyz, seg.scene_nrm, seg.delta); e.set_model(1, seg.model_xyz, seg.model_nrm)
e.generate_pcs(1, seed=3, max_hyp=20000)
e.score_generated(1, "weighted")
top = e.topk(1, 64)
poses = e.centred_to_pose(1, top["T"])
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, iters, en = e.tricp(1, seg.scene_xyz, poses, trim=0.5, ratio=0.99, max_iter=100)
    dt = time.perf_counter() - t0
    print('tricp ms', round(dt * 1e3, 3), 'poses/s', round(len(poses) / dt), 'iters mean', iters.mean(), 'max', iters.max())
