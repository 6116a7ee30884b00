#!/usr/bin/env python
"""bench.py -- cost-volume pixel*labels/s of the plane+parallax hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference's CPU path

One step = one pass of the hot path over one batch (``--batch``, default 16) of synthetic Sintel-shape
triplets that are already resident in HBM; per triplet: flow -> residual flow -> parallax -> mu, B ->
normed structures -> label range -> channels of the three frames -> per-label data cost of both
neighbour frames (51 labels, min / argmin) -> structure -> flow.  Work units = 2 neighbour frames x h x w
x 51 labels per triplet.  At N > 1 every rank owns its own triplets
(triplet-parallel, mrflow_dataset.py-style; no data-path collective): weak scaling.  The same line carries
``bands_4k``: ONE 4K frame sharded by row bands with halos over the N GPUs (strong scaling; NCCL only for
the small statistics collectives), at N = 1 the whole 4K frame on one GPU, the base of that speed-up.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_LABELS = 51
# Label range rule: the orphaned reference stage zeroes the pixels with rigidity > 0 before taking the
# 1.5x[min, max] range (pipeline/build_cost_volume.py:159), which on an all-rigid triplet collapses
# the 51 labels to zero parallax; the benchmark uses the rule the code evidently intended (range over
# the rigid pixels) so that the labels span the scene's real parallax range.
RANGE_MASK = "rigid_only"
METRIC = "cost_volume_pixel_labels_per_s"
UNIT = "pixel*labels/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="sintel", help="sintel | kitti | 4k | HxW")
    ap.add_argument("--interp", default="cubic", choices=["cubic", "linear"])
    ap.add_argument("--variant", default="staged", choices=["staged", "gather", "staged_occ2", "gather_occ2"])
    ap.add_argument("--batch", type=int, default=16, help="triplets per step (a step = one pass of the path over one batch)")
    ap.add_argument("--no-bands", action="store_true", help="skip the 4K row-band record (bands_4k)")
    ap.add_argument("--bands-timeout", type=int, default=240, help="seconds after which the line is printed without bands_4k")
    ap.add_argument("--ref-labels", type=int, default=None, help="--impl reference: labels per step (default: all 51)")
    ap.add_argument("--no-pair", action="store_true", help="one cost-volume launch per neighbour frame instead of one per triplet")
    ap.add_argument("--no-priority", action="store_true", help="pre-pass on a normal-priority stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the pre-pass eagerly instead of replaying its CUDA graph")
    ap.add_argument("--pipeline", type=int, default=3, help="triplets in flight per GPU (dataset-style software pipelining)")
    return ap.parse_args()


def shape_of(s):
    from mrflow_b200 import synth
    if s in synth.SHAPES:
        return synth.SHAPES[s]
    h, w = s.lower().split("x")
    return int(h), int(w)


def workload_name(args, hw):
    return "%s_%dx%d_L%d_F2_%s" % (args.shape, hw[1], hw[0], N_LABELS, args.interp)


def bench_config(args, hw):
    """the same dict in both arms (the driver compares them)"""
    vol_mb = 2 * N_LABELS * hw[0] * hw[1] * 4 / 1e6
    l2 = ("GPU arm: L2 flushed (256 MiB fill) before the timed region and before every single-triplet step; a step cycles "
          "over %d resident triplets and writes %d x %.0f MB (> 126 MB L2).  CPU arm: not applicable"
          % (max(1, args.pipeline), max(1, args.batch), vol_mb))
    return {"workload": workload_name(args, hw), "labels": N_LABELS, "frames": 2, "interp": args.interp, "l2": l2}


def algorithmic_bytes_per_launch(h, w, L=N_LABELS):
    """one launch computes BOTH neighbour frames of a triplet; every input read once and
    every output written once: per frame the cost store L*4, min/argmin 4, neighbour channels 3*4, masks 2, plus
    the reference-frame channels 3*4 once = h*w*(2*(4L + 18) + 12) = 456 B/px at L = 51 (4.47 B per pixel*label)."""
    return h * w * (2 * (L * 4 + 4 + 12 + 2) + 12)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(t, ch, state, label_idx, interp="cubic"):
    """One bounded sample of the workload on the host: the oracle's cost volume (reference sampler
    cv2.remap + reference geometry + Lorentzian) for ``label_idx`` of the 51 labels on both frames.
    Returns (seconds, pixel*labels done)."""
    from oracle import hotpath as hp
    h, w = t["shape"]
    noocc = [np.zeros((h, w), bool)] * 2
    t0 = time.perf_counter()
    hp.cost_volume(ch, t["rigidity"], noocc, t["homographies"], state["B"], state["mu"], state["q"], state["labels"],
                   kind="lorentzian", sigma=1.0, labels=label_idx, mode=interp)
    return time.perf_counter() - t0, 2 * h * w * len(label_idx)


def cpu_front_and_tail(t):
    """The non-cost-volume part of a step on the host (structure pre-pass, label range, channels,
    structure -> flow).  Returns (seconds, state)."""
    from oracle import hotpath as hp
    h, w = t["shape"]
    y, x = np.mgrid[:h, :w]
    t0 = time.perf_counter()
    par, q_new, mu_ar, dn = [], [], [], []
    rig = t["rigidity"]
    for f in range(2):
        ur, vr = hp.get_residual_flow(x, y, t["flow"][f][0], t["flow"][f][1], t["homographies"][f])
        qn = hp.correct_epipole(ur, vr, t["epipoles"][f], rig > 0)
        p, _, d = hp.flow2parallax(ur, vr, qn)
        par.append(p); q_new.append(qn); mu_ar.append(d.mean()); dn.append(d / mu_ar[-1])
    pv = np.logical_and(np.logical_and(np.abs(par[0]) > 0.1, np.abs(par[1]) > 0.1), rig > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        target = par[1][pv] / (1.0 * (par[1][pv] / mu_ar[1] - dn[1][pv]))
        template = mu_ar[1] / mu_ar[0] * par[0][pv] / (par[0][pv] / mu_ar[0] - dn[0][pv])
        B_ar = [np.median(template / target), 1.0]
    S = [hp.parallax2normedstructure(par[f], q_new[f], B_ar[f], mu_ar[f]) for f in range(2)]
    labels = hp.structure_range(S, rig, q_new, range_mask=RANGE_MASK)
    ch = [hp.prepare_channels(I) for I in t["images"]]
    hp.combine_flow(S[1], t["flow"][1], (rig > 0).astype(np.float64), q_new, mu_ar, t["homographies"], B_ar)
    dt = time.perf_counter() - t0
    return dt, dict(B=B_ar, mu=mu_ar, q=q_new, labels=labels, S=S), ch


def host_threads():
    try:
        import cv2
        return int(cv2.getNumThreads())
    except Exception:   # noqa: BLE001
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mrflow_b200 import synth
    hw = shape_of(args.shape)
    steps = 8 if args.steps is None else args.steps
    warm = 1 if args.warmup is None else args.warmup
    t = synth.make_triplet(hw, seed=1001)
    t_rest, state, ch = cpu_front_and_tail(t)
    # every step is the WHOLE workload of one triplet on the host: all 51 labels of both neighbour frames (about 3 s
    # on 16 cores) plus the pre-pass / channels / structure->flow part (timed once, identical every step); nothing is
    # pro-rated.  --ref-labels N bounds the sample to N evenly spaced labels (then the rest is scaled by N/51).
    n_lab = N_LABELS if args.ref_labels is None else max(1, min(N_LABELS, args.ref_labels))
    idx = sorted(set(int(round(v)) for v in np.linspace(0, N_LABELS - 1, n_lab)))
    for _ in range(warm):
        cpu_reference_step(t, ch, state, idx, args.interp)
    tot_s, tot_u = 0.0, 0
    for _ in range(steps):
        dt, units = cpu_reference_step(t, ch, state, idx, args.interp)
        tot_s += dt + t_rest * len(idx) / N_LABELS
        tot_u += units
    value = tot_u / tot_s
    cores = host_threads()
    sample = "%d of %d labels x 2 frames per step on the full %dx%d triplet; structure pre-pass / channels / " \
             "structure->flow timed once (%.3f s) and added to every step%s" % (
                 len(idx), N_LABELS, hw[1], hw[0], t_rest, "" if len(idx) == N_LABELS else " in proportion")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * tot_s / max(steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, hw),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def parity_against_oracle(t, gpu_cost, state, planes, interp="cubic"):
    """The checker (cpu_baseline leg): label planes of the GPU's volumes against the CPU oracle evaluated at the
    GPU's own whole-frame scalars.  ``gpu_cost``: the two device volumes; ``planes``: label indices."""
    from oracle import hotpath as hp
    h, w = t["shape"]
    ch = [hp.prepare_channels(I) for I in t["images"]]
    noocc = [np.zeros((h, w), bool)] * 2
    t0 = time.perf_counter()
    want, _, _ = hp.cost_volume(ch, t["rigidity"], noocc, t["homographies"], state["B"], state["mu"], state["q"], state["labels"],
                                kind="lorentzian", sigma=1.0, labels=planes, mode=interp)
    dt = time.perf_counter() - t0
    worst, exact, n = 0.0, 0, 0
    for f in range(2):
        got = gpu_cost[f][planes].cpu().numpy().astype(np.float64)
        ref = want[f].astype(np.float64)
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2)
        worst = max(worst, float(err.max()))
        exact += int((got == ref).sum()); n += got.size
    return {"max_rel_err": worst, "frac_bit_exact": exact / n, "labels_checked": len(planes), "frames": 2,
            "against": "CPU oracle (oracle/hotpath.py) at the GPU's mu / B / labels, same triplet as the timed one",
            "tolerance": 1e-5}, dt, 2 * h * w * len(planes)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mrflow_b200 import device, planeparallax as pp, synth
    from mrflow_b200.pipeline import build_cost_volume as bcv
    from mrflow_b200.pipeline import structure2flow as s2f

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; mrflow_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = 20 if args.steps is None else args.steps
    warm = 5 if args.warmup is None else max(args.warmup, 3)
    hw = shape_of(args.shape)
    h, w = hw
    T = max(1, args.batch)                       # triplets per step

    t = synth.make_triplet(hw, seed=1001 + rank)
    rig_host = t["rigidity"] > 0
    band = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None)
    rigid_mask = torch.from_numpy(rig_host.astype(np.uint8)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    # ``--pipeline P``: P distinct triplets resident per GPU, each with its own stream; consecutive triplets of a
    # step alternate between them, so the short pre-pass launches of one overlap the cost-volume kernels of
    # another (what a dataset run over many triplets does).  A step = one pass of the path over a batch of
    # ``--batch`` triplets (default 16: ~10 ms, so that 20 timed steps make a 0.2 s region).
    P = max(1, args.pipeline)
    slots = []
    host_triplets = []
    for s_i in range(P):
        ts = t if s_i == 0 else synth.make_triplet(hw, seed=2001 + 17 * s_i + rank)
        host_triplets.append(ts)
        bs = band if s_i == 0 else pp.Band.from_host(ts["images"], ts["flow"], ts["rigidity"], None)
        ms = rigid_mask if s_i == 0 else torch.from_numpy((ts["rigidity"] > 0).astype(np.uint8)).cuda()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            st = bcv.ResidentStep(bs, ts["homographies"], ts["epipoles"], ms, rho="lorentzian", sigma=1.0,
                                  interp=args.interp, range_mask=RANGE_MASK, variant=args.variant,
                                  use_graph=not args.no_graph, pair=not args.no_pair, priority=not args.no_priority)
        stream.synchronize()
        slots.append((st, stream))
    stepper = slots[0][0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k_run = 0
    for _ in range(warm):
        for _ in range(T):
            st, stream = slots[k_run % P]
            with torch.cuda.stream(stream):
                st.run()
            k_run += 1
    barrier()
    device.reset_launch_count()
    main = torch.cuda.current_stream()
    ev_batch, ev_single = [], []
    n_single = max(steps, 10)
    with ClockSampler(local) as clocks:
        # (1) THE timed region: K steps, each one batch of T triplets, P in flight
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _, stream in slots:
            stream.wait_stream(main)
        for _ in range(steps * T):
            st, stream = slots[k_run % P]
            with torch.cuda.stream(stream):
                st.run(ev_batch)
            k_run += 1
        for _, stream in slots:
            main.wait_stream(stream)
        e1.record(main)
        e1.synchronize()
        total_ms = e0.elapsed_time(e1)
        # (2) one triplet at a time, L2 flushed before each: the latency of the path
        single_ms = 0.0
        for k in range(n_single):
            flush.fill_(1)
            slots[0][1].wait_stream(main)
            with torch.cuda.stream(slots[0][1]):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                slots[0][0].run(ev_single)
                a1.record()
            a1.synchronize()
            single_ms += a0.elapsed_time(a1)
    n_runs = steps * T + n_single
    launches = device.launch_count() + n_runs * stepper.launches_in_graph     # replayed launches are not re-counted
    barrier()
    tmax = torch.tensor([total_ms, single_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max, single_ms_max = float(tmax.cpu()[0]), float(tmax.cpu()[1])
    units_per_triplet = 2 * h * w * N_LABELS
    value = world * units_per_triplet * T * steps / (total_ms_max * 1e-3)

    # dominant kernel: k_cost_volume, ONE launch per triplet (both neighbour frames, grid.z = frame).  Its duration
    # is taken from CUDA events on the launching stream in the single-triplet region, where the kernel has the
    # GPU to itself (L2 flushed before each triplet): that is the kernel's own roofline.  In the region of `value`
    # three triplets are in flight and a launch shares the SMs with the other triplets' pre-pass kernels, so events
    # round it there measure a stretched interval; what that region does give rigorously is a lower bound: the
    # whole step time attributed to the kernel (value_region.frac_lower_bound).
    k_avg_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_batch]))
    k_single_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_single]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    nbytes = algorithmic_bytes_per_launch(h, w)
    achieved = nbytes / (k_single_ms * 1e-3) / 1e9
    traffic, traffic_src, inst = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(workload_name(args, hw))
        traffic_src = tj.get("_captures", {}).get(workload_name(args, hw)) or tj.get("_capture")
        inst = tj.get("_inst_executed", {}).get(workload_name(args, hw))
    except Exception:   # noqa: BLE001
        pass
    ms_per_triplet = total_ms_max / (steps * T)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "k_cost_volume (both neighbour frames of a triplet in one launch)", "kernel_ms": k_single_ms,
                "launches_per_triplet": 1, "algorithmic_bytes_per_launch": nbytes,
                "timed_in": "single-triplet region (the kernel alone on the GPU, L2 flushed; CUDA events on the launching stream)",
                "kernel_share_of_step": k_single_ms / (single_ms_max / n_single),
                "value_region": {"ms_per_triplet": ms_per_triplet, "kernel_ms_events": k_avg_ms,
                                 "frac_lower_bound": nbytes / (ms_per_triplet * 1e-3) / 1e9 / peak,
                                 "note": "3 triplets in flight: events round a launch include time shared with other "
                                         "triplets' kernels; the bound attributes the whole step to the kernel"},
                "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"}
    clk = clocks.summary().get("sm_mhz")
    if inst and clk:
        # the kernel is bound by instruction issue and shared-memory bandwidth, not by HBM: warp instructions
        # executed per launch (from the committed ncu capture of this kernel) over the issue slots of the launch
        roofline["issue_frac"] = inst / (148 * 4 * clk * 1e6 * k_single_ms * 1e-3)
        roofline["issue_frac_note"] = "smsp__inst_executed.sum of the capture in traffic_source / (148 SMs x 4 schedulers x SM clock x kernel_ms)"

    # ---- end to end through the public host API (numpy in pinned memory -> numpy out)
    e2e = None
    if not args.no_e2e:
        from mrflow_b200 import dataset

        def pinned(a):
            tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return tt.numpy()
        hosts = []
        for ts in host_triplets:
            hosts.append(dict(images=[pinned(I) for I in ts["images"]],
                              flow=[[pinned(ts["flow"][f][0]), pinned(ts["flow"][f][1])] for f in range(2)],
                              rigidity=pinned(ts["rigidity"]), homographies=ts["homographies"], epipoles=ts["epipoles"]))
        h0 = hosts[0]
        h2d = sum(a.nbytes for a in h0["images"]) + sum(a.nbytes for f in h0["flow"] for a in f) + h0["rigidity"].nbytes
        # the public dataset API: P triplets in flight, every triplet uploads its inputs from pinned host memory
        # and downloads both cost volumes, both structures, the flow and the scalars into pinned host memory
        stream_api = dataset.TripletStream(depth=P, interp=args.interp, variant=args.variant, range_mask=RANGE_MASK)
        n_e2e = steps * T

        def e2e_run(n):
            last = None
            for k in range(n):
                if len(stream_api.pending) == len(stream_api.slots):
                    last = stream_api.collect()
                hk = hosts[k % P]
                stream_api.submit(hk["images"], hk["flow"], hk["rigidity"], hk["homographies"], hk["epipoles"])
            while stream_api.pending:
                last = stream_api.collect()
            return last
        out = e2e_run(2 * P)
        d2h = sum(a.nbytes for a in out["cost"]) + sum(a.nbytes for a in out["structure"]) + out["u"].nbytes + out["v"].nbytes
        d2h += 8 * (16 + 256)
        barrier()
        t0 = time.perf_counter()
        e2e_run(n_e2e)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * units_per_triplet * n_e2e / float(tt.cpu()[0]), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d) * T, "d2h_bytes_per_step": int(d2h) * T, "steps": steps, "triplets_per_step": T,
               "triplets_in_flight": P, "ms_per_triplet": 1e3 * float(tt.cpu()[0]) / n_e2e,
               "h2d_bytes_per_triplet": int(h2d), "d2h_bytes_per_triplet": int(d2h),
               "pcie_gbs_achieved": world * (h2d + d2h) * n_e2e / float(tt.cpu()[0]) / 1e9,
               "api": "mrflow_b200.dataset.TripletStream (build_cost_volume + combine_flow per triplet, host numpy in/out)"}
        # one call at a time through the reference-signature functions, for comparison
        h1 = hosts[0]
        for _ in range(6):          # the pinned result buffers settle after a few calls (two result sets alive at once)
            o1 = bcv.build_cost_volume(h1["images"], h1["flow"], h1["rigidity"], h1["homographies"], h1["epipoles"], None,
                                       interp=args.interp, variant=args.variant, range_mask=RANGE_MASK)
            s2f.combine_flow(o1[1][1], h1["flow"][1], h1["rigidity"], o1[4], o1[2], h1["homographies"], o1[3])
        torch.cuda.synchronize()
        n1 = 15
        times = []
        for _ in range(n1):
            t0 = time.perf_counter()
            o1 = bcv.build_cost_volume(h1["images"], h1["flow"], h1["rigidity"], h1["homographies"], h1["epipoles"], None,
                                       interp=args.interp, variant=args.variant, range_mask=RANGE_MASK)
            s2f.combine_flow(o1[1][1], h1["flow"][1], h1["rigidity"], o1[4], o1[2], h1["homographies"], o1[3])
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        e2e["single_call"] = {"value": units_per_triplet / float(np.median(times)), "calls": n1,
                              "ms_per_call_median": float(np.median(times)) * 1e3, "ms_per_call_max": float(max(times)) * 1e3,
                              "ms_per_call_all": [round(x * 1e3, 3) for x in times],
                              "api": "mrflow_b200.pipeline.build_cost_volume + combine_flow, one triplet at a time"}

    # ---- the CPU leg (rank 0): the oracle as the checker of the volumes just timed and, at N = 1, as the baseline
    cpu_baseline, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        mu_g, B_g, labels_g = pp.state_to_host(stepper.buf.state)
        state_g = dict(mu=mu_g, B=B_g, labels=labels_g, q=stepper.q)
        if world == 1:
            t_rest, _, _ = cpu_front_and_tail(t)
            idx = list(range(N_LABELS))                      # the whole workload once: ~5-10 s of host time
            parity, dt, units = parity_against_oracle(t, stepper.cost, state_g, idx, args.interp)
            cpu_baseline = {"value": units / (dt + t_rest * len(idx) / N_LABELS), "unit": UNIT, "cores": host_threads(),
                            "kind": "port", "host_cpus": os.cpu_count(),
                            "sample": "%d of %d labels x 2 frames of the same triplet, 1 pass (%.1f s) + pre-pass / channels / "
                                      "structure->flow (%.2f s)" % (len(idx), N_LABELS, dt, t_rest)}
        else:
            parity, _, _ = parity_against_oracle(t, stepper.cost, state_g, [0, 25, 50], args.interp)

    stepper_launches = stepper.launches_in_graph
    cuda_graph = stepper.graph is not None

    printed = []

    def emit(bands_rec):
        printed.append(1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": total_ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 geometry / f32 sampling", "data": "synthetic",
                "config": bench_config(args, hw),
                "details": {"triplets_per_step": T, "ms_per_triplet": ms_per_triplet, "variant": args.variant,
                            "range_mask": RANGE_MASK,
                            "cuda_graph": "pre-pass + channels (%d launches) replayed as one graph" % stepper_launches
                            if cuda_graph else "off", "triplets_in_flight": P, "parallelism": "triplets x%d" % world,
                            },
                "single_triplet": {"ms_per_step": single_ms_max / n_single, "steps": n_single,
                                   "value": world * units_per_triplet * n_single / (single_ms_max * 1e-3),
                                   "note": "one triplet at a time, L2 flushed before every step (path latency)"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "parity": parity, "bands_4k": bands_rec}
        print(json.dumps(line), flush=True)

    # ---- BASELINE.json configs[3]: one 4K frame in row bands over the N GPUs (strong scaling); at N = 1 the base
    bands, dog = None, None
    if not args.no_bands:
        slots.clear()
        stepper = None
        torch.cuda.empty_cache()
        from mrflow_b200 import sharding
        # The headline line must come out whatever happens in this add-on record: an exception becomes an "error"
        # entry, and a watchdog prints the line without the record if the banded run does not return in time (a
        # lost rank would otherwise leave the others in a collective).
        def _bail():
            if rank == 0 and not printed:
                emit({"error": "bands_4k did not finish within %d s" % args.bands_timeout})
            os._exit(0)
        dog = threading.Timer(args.bands_timeout, _bail)
        dog.daemon = True
        dog.start()
        try:
            bands = sharding.bench_bands(synth.SHAPES["4k"], max(5, min(steps, 20)), 3, interp=args.interp, variant=args.variant)
        except Exception as e:                              # noqa: BLE001
            bands = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        emit(bands)
    if world > 1:
        dist.destroy_process_group()          # still under the watchdog: a rank that failed above may not join
    if dog is not None:
        dog.cancel()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
