"""Small, fast exercise of every kernel family for compute-sanitizer memcheck (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mrflow_b200 import device, planeparallax as pp, synth
from mrflow_b200.pipeline import build_cost_volume as bcv

t = synth.make_triplet((72, 96), seed=8, epipole="inside", n_waves=4)
t["rigidity"][10:20, 30:50] = 0.0
band = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None)
mask = torch.from_numpy((t["rigidity"] > 0).astype(np.uint8)).cuda()
for interp in ("cubic", "linear"):
    for variant in ("staged", "gather"):
        r = bcv.cost_volume_pass_resident(band, t["homographies"], t["epipoles"], mask, interp=interp, variant=variant,
                                          range_mask="rigid_only")
a = bcv.cost_volume_pass(band, t["homographies"], t["epipoles"], mask, pp.LocalStats(), range_mask="rigid_only", layout="HWL_I32")
# a row band with a halo that is too small (exercises the halo_miss path)
b2 = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None, rows=(0, 48), halo=1)
miss = torch.zeros(1, dtype=torch.int32, device="cuda")
bcv.cost_volume_pass(b2, t["homographies"], t["epipoles"], mask[0:48].contiguous(), pp.LocalStats(), range_mask="rigid_only", halo_miss=miss)
ch = pp.channels(band)
device.data_term(torch.from_numpy(t["A_true"]).cuda(), ch[0], ch[0].clone(), (mask == 0).to(torch.uint8), torch.zeros_like(mask),
                 t["homographies"][1], 0.4, t["mu"][1], t["epipoles"][1])
fl = [[torch.from_numpy(t["flow"][f][k]).cuda() for k in range(2)] for f in range(2)]
device.flow_consistency(fl, fl, 2.0)
device.image_weights(torch.from_numpy(t["images"][1]).cuda())
# round 2: both forms of the pre-pass, the auto-sigma cost volume, the device-resident structure stage, the
# pipelined host call, the five-pass select on its own
rigid = mask
nonrigid = (mask == 0).to(torch.uint8)
for fused in (False, True):
    buf = device.FrontBuffers(72, 96)
    device.cost_volume_front(band.flow, rigid, nonrigid, t["homographies"], t["epipoles"], 72, buf, fused=fused)
bcv.build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], None, rho="lorentzian", sigma=None,
                      range_mask="rigid_only")
bcv.build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], None, range_mask="rigid_only")
bcv.build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], None, range_mask="rigid_only",
                      layout="HWL_I32")
from mrflow_b200.pipeline import compute_structure as cs
import argparse
cs.compute_structure(t["images"], t["flow"], t["rigidity"], [np.zeros((72, 96), bool)] * 2, [H.copy() for H in t["homographies"]],
                     t["epipoles"], argparse.Namespace(nonlinear_structure_initialization=0),
                     infer_mrf=lambda I, un, lam: (un[:, :, 1] > un[:, :, 0]).astype(np.int32))
for n in (0, 1, 5000):
    device.median(torch.randn(max(n, 1) + 3, dtype=torch.float64, device="cuda"), n)
torch.cuda.synchronize()
print("sanitize_small ok, halo misses:", int(miss.cpu()[0]))
