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
torch.cuda.synchronize()
print("sanitize_small ok, halo misses:", int(miss.cpu()[0]))
