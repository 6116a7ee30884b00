#!/usr/bin/env python
"""Timings of the rows next to the cost volume at Sintel size: device time with
CUDA events (inputs resident in HBM, L2 flushed between iterations) beside the CPU oracle.  One JSON line
per row.  Run on the GPU box: `python tools/bench_rows.py > gpurun_out/rows.jsonl`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from mrflow_b200 import device, synth
    from oracle import hotpath as hp
    t = synth.make_triplet("sintel", seed=1001)
    h, w = t["shape"]
    cu = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    ch = [hp.prepare_channels(I) for I in t["images"]]
    rs = np.random.RandomState(0)
    import cv2
    back = [[(-t["flow"][f][k] + cv2.GaussianBlur(rs.standard_normal((h, w)).astype(np.float32), (0, 0), 4.0) * 4).astype(np.float32)
             for k in range(2)] for f in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fl = [[cu(t["flow"][f][0]), cu(t["flow"][f][1])] for f in range(2)]
    bf = [[cu(back[f][0]), cu(back[f][1])] for f in range(2)]
    I1d, I2d = cu(ch[1]), cu(ch[2])
    Ad = cu(t["A_true"])
    zero = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    imgd = cu(t["images"][1])

    def gpu_time(fn, n=10):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(n):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            tot += a.elapsed_time(b)
        return tot / n

    def cpu_time(fn, n=2):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3

    noocc = np.zeros((h, w))
    rows = [
        ("flow_consistency (rigidity/occlusions.py:10-74)", lambda: device.flow_consistency(fl, bf, 2.0),
         lambda: hp.compute_occlusions_from_consistency(t["flow"], back, 2.0)),
        ("image_weights (rigidity/inference.py:140-167)", lambda: device.image_weights(imgd), lambda: hp.get_image_weights(t["images"][1])),
        ("derivs_through_H (interpolate_through_H.py:53-97)", lambda: device.derivs_through_H(I2d, t["homographies"][1]),
         lambda: hp.compute_derivs_through_H(ch[2], t["homographies"][1])),
        ("computeDataTerm (optimize_structure_single_scale.py:186-305)",
         lambda: device.data_term(Ad, I1d, I2d, zero, zero, t["homographies"][1], t["B"][1], t["mu"][1], t["epipoles"][1]),
         lambda: hp.compute_data_term(t["A_true"], ch[1], ch[2], np.ones((h, w)), noocc, t["homographies"][1], t["B"][1],
                                      t["mu"][1], t["epipoles"][1])),
    ]
    for name, g, c in rows:
        gm, cm = gpu_time(g), cpu_time(c)
        print(json.dumps({"row": name, "shape": [h, w], "gpu_ms": gm, "cpu_oracle_ms": cm, "cpu_cores": os.cpu_count(),
                          "pixels_per_s_gpu": h * w / (gm * 1e-3), "speedup": cm / gm}))


if __name__ == "__main__":
    main()
