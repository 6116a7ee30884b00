#!/usr/bin/env python
"""Time k_cost_volume variants against each other on one GPU (CUDA events, L2 flushed before every launch)
and check that every variant produces the same bits as the first one.

    python tools/bench_kernel.py [--shape sintel] [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from mrflow_b200 import _lib, device, planeparallax as pp, synth
    from mrflow_b200.pipeline import build_cost_volume as bcv
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="sintel")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--variants", default="")
    ap.add_argument("--labels", type=int, default=51, help="labels evaluated per launch (the first n of the triplet's 51)")
    args = ap.parse_args()
    hw = synth.SHAPES[args.shape]
    h, w = hw
    t = synth.make_triplet(hw, seed=1001) if args.shape != "4k" else None
    if t is None:
        td = synth.make_triplet_device(hw, seed=1001, n_waves=16)
        from mrflow_b200 import sharding
        band = sharding.device_band(td, (0, h), 0)
        rigid = (td["rigidity"] > 0).to(torch.uint8)
        H_ar, q_ar = td["homographies"], td["epipoles"]
    else:
        band = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None)
        rigid = torch.from_numpy((t["rigidity"] > 0).astype(np.uint8)).cuda()
        H_ar, q_ar = t["homographies"], t["epipoles"]
    nonrigid = (rigid == 0).to(torch.uint8)
    buf = device.FrontBuffers(h, w)
    device.cost_volume_front(band.flow, rigid, nonrigid, H_ar, q_ar, h, buf)
    ref64, tex = pp.channels(band)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = torch.empty((51, h, w), dtype=torch.float32, device="cuda")
    res = []
    base = {}
    names = args.variants.split(",") if args.variants else None
    for interp in ("cubic", "linear"):
        for variant in sorted(_lib.VARIANT, key=lambda k: _lib.VARIANT[k]):
            if names and variant not in names:
                continue
            for f in (1,):
                def run():
                    return device.cost_volume(ref64, tex[f], None, H_ar[f], q_ar[f], 0.0, 0.0, nonrigid=nonrigid, interp=interp,
                                              variant=variant, out=out, state=buf.state, frame=f, n_labels=args.labels)
                for _ in range(3):
                    _, mn, am = run()
                torch.cuda.synchronize()
                key = (interp, f)
                if key not in base:
                    base[key] = (out.clone(), mn.clone(), am.clone())
                same = bool(torch.equal(out, base[key][0]) and torch.equal(mn, base[key][1]) and torch.equal(am, base[key][2]))
                ts = []
                for _ in range(args.reps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                res.append(dict(shape=args.shape, interp=interp, variant=variant, frame=f, labels=args.labels, ms_median=float(np.median(ts)),
                                ms_min=float(min(ts)), same_bits_as_first=same))
                print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
