#!/usr/bin/env python
"""The reference's dense chain after its inputs are loaded (mrflow.py:98-140) at Sintel size, with the
reference's DEFAULT parameters -- compute_structure -> optimize (variational refinement, 5 outer iterations,
parameters.py:38-47) -> combine_flow -- through the drop-in modules on the GPU (host numpy in, host numpy out,
wall clock) beside the CPU oracle (the same chain as the reference runs it: numpy / OpenCV / scipy sparse +
MINRES).  The MRF solver inside compute_structure is the reference's own C++ code on both sides (oracle/_ref)
and is timed separately.  One JSON line.  `python tools/bench_chain.py > gpurun_out/chain.json`"""
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
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from mrflow_b200 import synth
    from mrflow_b200.pipeline import compute_structure as cs, optimize_variational_refinement as ovr, structure2flow as s2f
    from oracle import hotpath as hp, trws, variational as ov
    t = synth.make_triplet("sintel", seed=1003, flow_noise=0.03)
    h, w = t["shape"]
    rs = np.random.RandomState(9)
    occ = [rs.rand(h, w) < 0.03, rs.rand(h, w) < 0.03]
    mrf_s = [0.0]

    def timed_mrf(I, unaries, lambd):
        t0 = time.perf_counter()
        r = trws.infer_mrf(I, unaries, lambd)
        mrf_s[0] += time.perf_counter() - t0
        return r
    if not trws.available():
        timed_mrf = lambda I, unaries, lambd: (unaries[:, :, 1] > unaries[:, :, 0]).astype(np.int32)      # noqa: E731
    p = argparse.Namespace(debug_use_rigidity_gt=0, debug_save_frames=0, debug_compute_figure=0, rigidity_sigma_direction=1.0,
                           rigidity_weight_cnn=0.25, rigidity_sigma_structure=1.0, rigidity_use_structure=1, scale_structure=1,
                           nonlinear_structure_initialization=0, occlusion_reasoning=1, tempdir=".", override_optimization=0,
                           variational_lambda_1storder=0.5, variational_lambda_2ndorder=5.0, variational_lambda_consistency=0.0,
                           variational_N_outer=5, variational_N_inner=1, variational_scale_factor=0.5, variational_N_scales=1,
                           variational_last_scale=0, variational_dataterm_grayscale=1, variational_min_weight=0.0)

    def gpu_chain():
        mrf_s[0] = 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        S, Hs, mu, B, q, rig, oc = cs.compute_structure(t["images"], t["flow"], t["rigidity"], occ,
                                                        [H.copy() for H in t["homographies"]], t["epipoles"], p, infer_mrf=timed_mrf)
        t1 = time.perf_counter()
        A = ovr.optimize(t["images"], S, rig, q, Hs, B, mu, oc, p)[0]
        t2 = time.perf_counter()
        _, u, v = s2f.combine_flow(A, t["flow"][1], rig, q, mu, Hs, B, t["images"], p)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return dict(compute_structure_ms=(t1 - t0 - mrf_s[0]) * 1e3, mrf_solver_ms=mrf_s[0] * 1e3, optimize_ms=(t2 - t1) * 1e3,
                    combine_flow_ms=(t3 - t2) * 1e3, total_ms=(t3 - t0) * 1e3, total_without_mrf_solver_ms=(t3 - t0 - mrf_s[0]) * 1e3), (A, u, v)

    gpu_chain()                                                         # warm-up (allocator, module load)
    runs = [gpu_chain() for _ in range(a.repeat)]
    best = min(runs, key=lambda r: r[0]["total_ms"])
    # per-stage median over the repeats (the CPU solver inside compute_structure makes the totals noisy)
    med = {k: float(np.median([r[0][k] for r in runs])) for k in runs[0][0]}
    best = (med, best[1])
    out = {"chain": "compute_structure -> optimize(variational, N_outer=5) -> combine_flow (mrflow.py:98-140)", "shape": [h, w],
           "gpu": med, "gpu_note": "per-stage median of %d runs" % len(runs), "gpu_runs_total_ms": [r[0]["total_ms"] for r in runs]}
    if not a.no_cpu:
        mrf_s[0] = 0.0
        t0 = time.perf_counter()
        want = hp.compute_structure(t["images"], t["flow"], t["rigidity"], occ, [H.copy() for H in t["homographies"]],
                                    t["epipoles"], hp.Params(), infer_mrf=timed_mrf)
        t1 = time.perf_counter()
        import cv2
        S, mu, rig = want[0], want[2], want[5]
        ob = cv2.medianBlur(occ[0].astype("uint8"), ksize=3) > 0
        of = cv2.medianBlur(occ[1].astype("uint8"), ksize=3) > 0
        both = ob * of
        merged = hp.optimize_override(S, mu, occ)
        ob2, of2 = ob.copy(), of.copy(); ob2[both] = 0; of2[both] = 0
        A_o, _ = ov.optimize_structure_single_scale(merged, S[0] * mu[1] / mu[0], S[1], ob2, of2, both, rig,
                                                    [np.asarray(I, np.float64) for I in t["images"]], want[1], want[3], want[4], mu,
                                                    0.5, 5.0, 0.0, N_outer=5)
        t2 = time.perf_counter()
        _, wu, wv = hp.combine_flow(A_o, t["flow"][1], rig, want[4], mu, want[1], want[3])
        t3 = time.perf_counter()
        cpu = dict(compute_structure_ms=(t1 - t0 - mrf_s[0]) * 1e3, mrf_solver_ms=mrf_s[0] * 1e3, optimize_ms=(t2 - t1) * 1e3,
                   combine_flow_ms=(t3 - t2) * 1e3, total_ms=(t3 - t0) * 1e3, total_without_mrf_solver_ms=(t3 - t0 - mrf_s[0]) * 1e3)
        A, u, v = best[1]
        fin = np.isfinite(wu) & np.isfinite(u)
        out.update(cpu_oracle=cpu, cpu_cores=os.cpu_count(),
                   speedup_without_mrf_solver=cpu["total_without_mrf_solver_ms"] / best[0]["total_without_mrf_solver_ms"],
                   speedup_total=cpu["total_ms"] / best[0]["total_ms"],
                   max_abs_diff_structure=float(np.abs(A - A_o)[np.isfinite(A_o)].max()),
                   max_abs_diff_flow_px=float(max(np.abs(u - wu)[fin].max(), np.abs(v - wv)[fin].max())))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
