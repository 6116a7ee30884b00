#!/usr/bin/env python
"""Time the exact device-side median (mrf_select_median_dev, the five-pass select) and the whole pre-pass of a
Sintel-shape triplet with CUDA events (eager launches, warm and with L2 flushed).

    python tools/bench_select.py [--reps 50]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps, flush=None):
    import torch
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(min(ts))


def main():
    import torch
    from mrflow_b200 import _lib, device, planeparallax as pp, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    L = _lib.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n in (446464, 465750, 1 << 20):
        rs = np.random.RandomState(n)
        for kind in ("ratios", "zeros70"):
            x = rs.standard_normal(n) * 3.0
            if kind == "zeros70":
                x = np.abs(x); x[rs.rand(n) < 0.7] = 0.0
            keys = torch.from_numpy(x).cuda()
            cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
            out = torch.zeros(1, dtype=torch.float64, device="cuda")
            scr = torch.zeros(L.mrf_select_scratch_bytes(0), dtype=torch.uint8, device="cuda")
            row = dict(what="median", n=n, kind=kind)
            fn = lambda: _lib.check(L.mrf_select_median_dev(P(keys), P(cnt), n, P(out), P(scr), st), "median")
            row["us_median"], row["us_min"] = timed(fn, args.reps)
            row["value"] = float(out.cpu()[0])
            row["numpy"] = float(np.median(x))
            print(json.dumps(row), flush=True)
    # the whole pre-pass of a Sintel-shape triplet (both frames: parallax, mu, candidates, median, structures, labels)
    hw = synth.SHAPES["sintel"]
    h, w = hw
    t = synth.make_triplet(hw, seed=1001)
    band = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None)
    rigid = torch.from_numpy((t["rigidity"] > 0).astype(np.uint8)).cuda()
    nonrigid = (rigid == 0).to(torch.uint8)
    buf = device.FrontBuffers(h, w)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    row = dict(what="pre-pass", shape="sintel")
    fn = lambda: device.cost_volume_front(band.flow, rigid, nonrigid, t["homographies"], t["epipoles"], h, buf)
    row["us_median"], row["us_min"] = timed(fn, args.reps)
    row["cold_us_median"], _ = timed(fn, args.reps, flush)
    row["state"] = buf.state[:8].cpu().tolist()
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
