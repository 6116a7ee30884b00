#!/usr/bin/env python
"""One whole-frame step of the device-resident pass (pre-pass, channels, both cost volumes, structure -> flow)
at a chosen shape with eager launches -- the program to run under `ncu --metrics gpu__time_duration.sum` for a
per-kernel account of a step (each launch serialised and cold-cache there).

    python tools/band_profile.py [--shape 4k] [--steps 2]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from mrflow_b200 import sharding, synth
    from mrflow_b200.pipeline import build_cost_volume as bcv
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4k")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    hw = synth.SHAPES[args.shape] if args.shape in synth.SHAPES else tuple(int(v) for v in args.shape.split("x"))
    td = synth.make_triplet_device(hw, seed=1001, n_waves=8)
    band = sharding.device_band(td, (0, hw[0]), 0)
    rigid = (td["rigidity"] > 0).to(torch.uint8)
    st = bcv.ResidentStep(band, td["homographies"], td["epipoles"], rigid, range_mask="rigid_only", use_graph=False)
    torch.cuda.synchronize()
    print("PROFILE-START", flush=True)
    for _ in range(args.steps):
        st.run()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
