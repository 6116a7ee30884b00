#!/usr/bin/env python
"""One outer iteration of the variational refinement (optimize_structure_single_scale.py:419-650) at Sintel
size: GPU (device-resident, wall clock around a synchronised step, includes the host MINRES recurrences)
beside the CPU oracle (scipy sparse + MINRES, what the reference runs).  One JSON line.
`python tools/bench_variational.py [--shape sintel|small] [--outer 2] > gpurun_out/variational.jsonl`"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="sintel")
    ap.add_argument("--outer", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from mrflow_b200 import device, synth, variational
    from mrflow_b200._host import flag, up
    from oracle import variational as ov
    t = synth.make_triplet(a.shape if a.shape == "sintel" else (72, 96), seed=1001)
    h, w = t["shape"]
    rs = np.random.RandomState(0)
    import cv2
    A_true = np.nan_to_num(t["A_true"])
    A_init = A_true + cv2.GaussianBlur(rs.standard_normal((h, w)), (0, 0), 3.0) * 0.05
    rig = t["rigidity"] > 0.5
    none = np.zeros((h, w), bool)
    vin = dict(A_init=A_init, A_bwd_reference=A_true * (t["mu"][0] / t["mu"][1]), A_fwd_reference=A_true, occlusions_bwd=none,
               occlusions_fwd=none, occlusions_both=none, rigidity=rig, images=[np.asarray(I, np.float64) for I in t["images"]],
               H_ar=t["homographies"], B_ar=[float(b) for b in t["B"]], q_ar=[np.asarray(q, np.float64) for q in t["epipoles"]],
               mu_ar=[float(m) for m in t["mu"]])
    imgs = [up(I, np.float64) for I in vin["images"]]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = variational.Refiner(imgs, flag(none), flag(none), flag(rig), vin["H_ar"], vin["B_ar"], vin["q_ar"], vin["mu_ar"],
                              up(vin["A_bwd_reference"], np.float64), up(vin["A_fwd_reference"], np.float64), 0.5, 5.0, 0.0)
    A = up(A_init, np.float64).clone()
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    steps = []
    for _ in range(a.outer):
        device.reset_launch_count()
        t0 = time.perf_counter()
        e = ref.step(A)
        torch.cuda.synchronize()
        steps.append(dict(ms=(time.perf_counter() - t0) * 1e3, minres_iterations=ref.last["info"]["itn"], energy=e,
                          launches=device.launch_count()))
    out = {"row": "optimizeStructureSingleScale outer iteration (optimize_structure_single_scale.py:419-650)", "shape": [h, w],
           "gpu_setup_ms": setup_ms, "gpu_steps": steps}
    if not a.no_cpu:
        t0 = time.perf_counter()
        A_o, en, tr = ov.optimize_structure_single_scale(lambda_1storder=0.5, lambda_2ndorder=5.0, lambda_consistency=0.0,
                                                         N_outer=a.outer, return_trace=True, **vin)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        out.update(cpu_oracle_ms_total=cpu_ms, cpu_oracle_ms_per_outer=cpu_ms / a.outer, cpu_cores=os.cpu_count(),
                   cpu_energies=en, max_abs_diff_A=float(np.abs(A.cpu().numpy() - A_o).max()),
                   median_abs_diff_A=float(np.median(np.abs(A.cpu().numpy() - A_o))),
                   speedup_per_outer=cpu_ms / a.outer / np.mean([s["ms"] for s in steps]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
