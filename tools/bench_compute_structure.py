#!/usr/bin/env python
"""Host-API latency of mrflow_b200.pipeline.compute_structure at Sintel size (pinned numpy in, numpy out; the MRF
consumer replaced by a threshold so that only the stage itself is timed).  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from mrflow_b200 import synth
    from mrflow_b200.pipeline import compute_structure as cs
    t = synth.make_triplet("sintel", seed=1003, flow_noise=0.03)
    h, w = t["shape"]
    rs = np.random.RandomState(9)
    occ = [rs.rand(h, w) < 0.03, rs.rand(h, w) < 0.03]

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    images = [pin(I) for I in t["images"]]
    flow = [[pin(c) for c in f] for f in t["flow"]]
    rig = pin(t["rigidity"])
    occ = [pin(o) for o in occ]
    p = argparse.Namespace(nonlinear_structure_initialization=0)
    mrf = lambda I, un, lam: (un[:, :, 1] > un[:, :, 0]).astype(np.int32)      # noqa: E731
    times = []
    for k in range(25):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cs.compute_structure(images, flow, rig, occ, [H.copy() for H in t["homographies"]], t["epipoles"], p, infer_mrf=mrf)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    tm = np.array(times[5:]) * 1e3
    print(json.dumps({"api": "mrflow_b200.pipeline.compute_structure (Sintel shape, pinned host arrays in, numpy out, MRF "
                             "consumer stubbed)", "ms_median": float(np.median(tm)), "ms_min": float(tm.min()),
                      "ms_max": float(tm.max()), "calls": len(tm)}))


if __name__ == "__main__":
    main()
