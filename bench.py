#!/usr/bin/env python
"""bench.py -- cost-volume pixel*labels/s of the plane+parallax hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference's CPU path

One step = one pass of the hot path over one synthetic triplet that is already resident in HBM:
flow -> residual flow -> parallax -> mu, B -> normed structures -> label range -> channels of the
three frames -> per-label data cost of both neighbour frames (51 labels, min / argmin) -> structure ->
flow.  Work units = 2 neighbour frames x h x w x 51 labels.  At N > 1 every
rank owns its own triplet (triplet-parallel, mrflow_dataset.py-style; no data-path collective): weak
scaling.  ``--mode bands`` runs ONE frame sharded by row bands with halos instead (strong scaling;
NCCL only for the tiny statistics all-reduces), used for the 4K configuration.

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
    ap.add_argument("--mode", default="triplets", choices=["triplets", "bands"])
    ap.add_argument("--band-allreduce", default="nccl", choices=["nccl", "peer"],
                    help="--mode bands: small all-reduces of the pre-pass through NCCL or through the peer-memory mailbox kernel")
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


def algorithmic_bytes_per_launch(h, w, L=N_LABELS):
    """per neighbour-frame launch, every input read once and every output
    written once: cost store L*4, min/argmin 4, neighbour channels 3*4, masks 2, plus half of the
    reference-frame channels (3*4, shared by the two launches) = 228 B/px at L = 51."""
    return h * w * (L * 4 + 4 + 12 + 2 + 6)


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
def cpu_reference_step(t, ch, state, label_idx):
    """One bounded sample of the workload on the host: the oracle's cost volume (reference sampler
    cv2.remap + reference geometry + Lorentzian) for ``label_idx`` of the 51 labels on both frames.
    Returns (seconds, pixel*labels done)."""
    from oracle import hotpath as hp
    h, w = t["shape"]
    noocc = [np.zeros((h, w), bool)] * 2
    t0 = time.perf_counter()
    hp.cost_volume(ch, t["rigidity"], noocc, t["homographies"], state["B"], state["mu"], state["q"], state["labels"],
                   kind="lorentzian", sigma=1.0, labels=label_idx)
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
    idx = [0, 17, 34, 50]
    for _ in range(warm):
        cpu_reference_step(t, ch, state, idx)
    tot_s, tot_u = 0.0, 0
    for _ in range(steps):
        dt, units = cpu_reference_step(t, ch, state, idx)
        tot_s += dt + t_rest * len(idx) / N_LABELS       # pro-rated share of the non-cost-volume work
        tot_u += units
    value = tot_u / tot_s
    cores = host_threads()
    sample = "%d of %d labels x 2 frames per step on the full %dx%d triplet; structure pre-pass / channels / " \
             "structure->flow timed once (%.3f s) and pro-rated" % (len(idx), N_LABELS, hw[1], hw[0], t_rest)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * tot_s / max(steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, hw), "labels": N_LABELS, "frames": 2, "interp": args.interp},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
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
    steps = 30 if args.steps is None else args.steps
    warm = 5 if args.warmup is None else max(args.warmup, 3)
    hw = shape_of(args.shape)
    h, w = hw

    if args.mode == "bands" and world > 1:
        from mrflow_b200 import sharding
        return sharding.bench_bands(args, hw, steps, warm, METRIC, UNIT, workload_name(args, hw), ClockSampler)

    t = synth.make_triplet(hw, seed=1001 + rank)
    rig_host = t["rigidity"] > 0
    band = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None)
    rigid_mask = torch.from_numpy(rig_host.astype(np.uint8)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2
    kernel_events = []

    # ``--pipeline P``: P distinct triplets resident per GPU, each with its own stream; consecutive steps
    # alternate between them, so the short pre-pass launches of one triplet overlap the cost-volume kernels
    # of the other (what a dataset run over many triplets does).  Every step still processes one whole
    # triplet; P = 1 is the plain back-to-back case.
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
                                  use_graph=not args.no_graph)
        stream.synchronize()
        slots.append((st, stream))
    stepper = slots[0][0]

    def step(k, events=None):
        st, stream = slots[k % P]
        with torch.cuda.stream(stream):
            return st.run(events)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(warm * P):
        step(k)
    barrier()
    device.reset_launch_count()
    main = torch.cuda.current_stream()
    with ClockSampler(local) as clocks:
        # (1) one triplet at a time: per-step CUDA events, L2 flushed before every step.  This is the
        #     latency of the path and the region in which the dominant kernel is timed for the roofline.
        single_ms = 0.0
        for k in range(steps):
            flush.fill_(1)
            slots[0][1].wait_stream(main)
            with torch.cuda.stream(slots[0][1]):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                slots[0][0].run(kernel_events)
                e1.record()
            e1.synchronize()
            single_ms += e0.elapsed_time(e1)
        total_ms = single_ms
        # (2) P triplets in flight (the dataset-run form): one timed region over K steps
        if P > 1:
            barrier()
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for _, stream in slots:
                stream.wait_stream(main)
            for k in range(steps):
                step(k)
            for _, stream in slots:
                main.wait_stream(stream)
            e1.record(main)
            e1.synchronize()
            total_ms = e0.elapsed_time(e1)
    n_runs = steps * (2 if P > 1 else 1)
    launches = device.launch_count() + n_runs * stepper.launches_in_graph     # replayed launches are not re-counted
    barrier()
    tmax = torch.tensor([total_ms, single_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max, single_ms_max = float(tmax.cpu()[0]), float(tmax.cpu()[1])
    units_per_step = 2 * h * w * N_LABELS
    value = world * units_per_step * steps / (total_ms_max * 1e-3)

    # dominant kernel: k_cost_volume (one launch per neighbour frame)
    kms = [a.elapsed_time(b) for a, b in kernel_events]
    k_avg_ms = float(np.mean(kms))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algorithmic_bytes_per_launch(h, w) / (k_avg_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload_name(args, hw))
    except Exception:   # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_cost_volume", "kernel_ms": k_avg_ms,
                "kernel_share_of_step": 2 * k_avg_ms * steps / single_ms, "timed_in": "single-triplet region",
                "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"}

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
        h2d += 2 * h * w                                       # rigid / non-rigid flags
        # the public dataset API: P triplets in flight, every step uploads its inputs from pinned host memory
        # and downloads both cost volumes, both structures, the flow and the scalars into pinned host memory
        stream_api = dataset.TripletStream(depth=P, interp=args.interp, variant=args.variant, range_mask=RANGE_MASK)
        n_e2e = max(6, min(steps, 20))

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
        e2e = {"value": world * units_per_step * n_e2e / float(tt.cpu()[0]), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e, "triplets_in_flight": P,
               "api": "mrflow_b200.dataset.TripletStream (build_cost_volume + combine_flow per triplet, host numpy in/out)"}
        # one call at a time through the reference-signature functions, for comparison
        h1 = hosts[0]
        for _ in range(3):          # the pinned result buffers of `down()` settle after three calls (two sets alive at once)
            o1 = bcv.build_cost_volume(h1["images"], h1["flow"], h1["rigidity"], h1["homographies"], h1["epipoles"], None,
                                       interp=args.interp, variant=args.variant, range_mask=RANGE_MASK)
        torch.cuda.synchronize()
        n1 = 9
        times = []
        for _ in range(n1):
            t0 = time.perf_counter()
            o1 = bcv.build_cost_volume(h1["images"], h1["flow"], h1["rigidity"], h1["homographies"], h1["epipoles"], None,
                                       interp=args.interp, variant=args.variant, range_mask=RANGE_MASK)
            s2f.combine_flow(o1[1][1], h1["flow"][1], h1["rigidity"], o1[4], o1[2], h1["homographies"], o1[3])
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        # median call: the pinned result buffers are re-allocated now and then, which stalls single calls for tens of ms
        e2e["single_call"] = {"value": units_per_step / float(np.median(times)), "calls": n1,
                              "ms_per_call_median": float(np.median(times)) * 1e3, "ms_per_call_max": float(max(times)) * 1e3,
                              "api": "mrflow_b200.pipeline.build_cost_volume + combine_flow, one triplet at a time"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t_rest, state, ch = cpu_front_and_tail(t)
        idx = list(range(N_LABELS))                          # the whole workload once: ~5-10 s of host time
        dt, units = cpu_reference_step(t, ch, state, idx)
        cpu_baseline = {"value": units / (dt + t_rest * len(idx) / N_LABELS), "unit": UNIT, "cores": host_threads(),
                        "kind": "port", "host_cpus": os.cpu_count(),
                        "sample": "%d of %d labels x 2 frames of the same triplet, 1 pass (%.1f s)" % (len(idx), N_LABELS, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": total_ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 geometry / f32 sampling", "data": "synthetic",
                "config": {"workload": workload_name(args, hw), "labels": N_LABELS, "frames": 2, "interp": args.interp,
                           "variant": args.variant, "range_mask": RANGE_MASK,
                           "cuda_graph": "pre-pass + channels (%d launches) replayed as one graph" % stepper.launches_in_graph
                           if stepper.graph is not None else "off", "triplets_in_flight": P, "parallelism": "triplets x%d" % world,
                           "l2": ("flushed between timed steps (256 MiB fill); outputs 182 MB > L2" if P == 1 else
                                  "flushed before the timed region; steps cycle over %d resident triplets and each writes "
                                  "182 MB (> L2)" % P)},
                "single_triplet": {"ms_per_step": single_ms_max / steps,
                                   "value": world * units_per_step * steps / (single_ms_max * 1e-3),
                                   "note": "one triplet at a time, L2 flushed before every step (path latency)"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
