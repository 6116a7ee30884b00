import argparse, cProfile, pstats, sys, os, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from mrflow_b200 import synth
from mrflow_b200.pipeline import compute_structure as cs
t = synth.make_triplet("sintel", seed=1003, flow_noise=0.03)
h, w = t["shape"]
rs = np.random.RandomState(9)
occ = [rs.rand(h, w) < 0.03, rs.rand(h, w) < 0.03]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
images = [pin(I) for I in t["images"]]
flow = [[pin(c) for c in f] for f in t["flow"]]
rig = pin(t["rigidity"]); occ = [pin(o) for o in occ]
p = argparse.Namespace(nonlinear_structure_initialization=0)
mrf = lambda I, un, lam: (un[:, :, 1] > un[:, :, 0]).astype(np.int32)
def call():
    cs.compute_structure(images, flow, rig, occ, [H.copy() for H in t["homographies"]], t["epipoles"], p, infer_mrf=mrf)
    torch.cuda.synchronize()
for _ in range(5): call()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): call()
pr.disable()
st = pstats.Stats(pr, stream=sys.stdout); st.sort_stats("cumulative").print_stats(35)
