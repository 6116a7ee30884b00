"""Host-side profile of stage 5 (optimize with the variational refinement) at Sintel size."""
import argparse, cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mrflow_b200 import synth
from mrflow_b200.pipeline import optimize_variational_refinement as ovr
t = synth.make_triplet("sintel", seed=1003, flow_noise=0.03)
h, w = t["shape"]
rs = np.random.RandomState(9)
occ = [rs.rand(h, w) < 0.03, rs.rand(h, w) < 0.03]
A = np.nan_to_num(t["A_true"])
S = [A * (t["mu"][0] / t["mu"][1]), A]
p = argparse.Namespace(override_optimization=0, occlusion_reasoning=1, tempdir=".", variational_lambda_1storder=0.5,
                       variational_lambda_2ndorder=5.0, variational_lambda_consistency=0.0, variational_N_outer=5, variational_N_inner=1,
                       variational_scale_factor=0.5, variational_N_scales=1, variational_last_scale=0, variational_dataterm_grayscale=1,
                       variational_min_weight=0.0)
call = lambda: ovr.optimize(t["images"], S, t["rigidity"] > 0.5, t["epipoles"], t["homographies"], t["B"], t["mu"], occ, p)
for _ in range(3): call()
torch.cuda.synchronize(); t0 = time.perf_counter(); call(); torch.cuda.synchronize(); print("optimize: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); call(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
