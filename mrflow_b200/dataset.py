"""Dataset runs: many independent triplets through ``build_cost_volume`` + ``combine_flow``
(the job of the reference's ``mrflow_dataset.py``, which spawns one ``mrflow.py`` process per triplet,
:121-169).

``TripletStream`` keeps a few triplets in flight on one GPU: each slot has its own CUDA stream and its own
pinned result buffers, so the host->device copy and the kernels of triplet i+1 overlap the device->host
copy of triplet i's 182 MB of cost volume (PCIe is full duplex; the download is the long pole).  Results
come back in submission order.  Across GPUs, give every rank its own share of the triplets -- there is no
collective ("triplet-parallel").
"""
import collections

import numpy as np
import torch

from . import device, planeparallax as pp
from ._host import flag, require_cuda
from .pipeline.build_cost_volume import N_RANGES, cost_volume_pass_resident


class _Slot:
    def __init__(self):
        self.stream = torch.cuda.Stream()
        self.done = torch.cuda.Event()
        self.host = {}          # name -> pinned tensor, reused across triplets
        self.keep = None        # device tensors that must outlive the asynchronous copies

    def pinned_like(self, name, t):
        buf = self.host.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self.host[name] = buf
        return buf


class TripletStream:
    """``submit(images, flow, rigidity, homographies, epipoles)`` enqueues one triplet; ``collect()`` returns
    the oldest finished result as a dict of numpy arrays (views of the slot's pinned buffers: valid until
    ``depth`` further triplets have been submitted) -- ``cost`` 2 x fp32 [51,h,w], ``structure`` 2 x fp64
    [h,w], ``mu``, ``B``, ``epipoles``, ``labels``, and the forward flow ``u``, ``v`` of ``combine_flow``."""

    def __init__(self, depth=2, rho="lorentzian", sigma=1.0, interp="cubic", range_mask="as_written", variant="staged"):
        require_cuda()
        self.slots = [_Slot() for _ in range(max(1, depth))]
        self.kw = dict(rho=rho, sigma=sigma, interp=interp, range_mask=range_mask, variant=variant)
        self.pending = collections.deque()
        self.n = 0

    def submit(self, images, flow, rigidity, homographies, epipoles):
        if len(self.pending) == len(self.slots):
            raise RuntimeError("all slots are in flight: collect() a result first")
        slot = self.slots[self.n % len(self.slots)]
        self.n += 1
        rigidity = np.asarray(rigidity)
        with torch.cuda.stream(slot.stream):
            band = pp.Band.from_host(images, flow, rigidity, None)
            r = cost_volume_pass_resident(band, homographies, epipoles, flag(rigidity > 0), **self.kw)
            buf = r["buffers"]
            # combine_flow (pipeline/structure2flow.py:57-82): forward structure -> flow, initial flow where
            # rigidity == 0; mu / B come from the device state
            u, v = device.structure_to_flow(buf.S[1], r["q"][1], 0.0, homographies[1], 0.0, nonrigid=flag(rigidity == 0),
                                            init_u=band.flow[1][0], init_v=band.flow[1][1], state=buf.state, frame=1)
            outs = {"cost0": r["cost"][0], "cost1": r["cost"][1], "S0": buf.S[0], "S1": buf.S[1], "u": u, "v": v,
                    "state": buf.state}
            host = {}
            for k, t in outs.items():
                hb = slot.pinned_like(k, t)
                hb.copy_(t, non_blocking=True)
                host[k] = hb
            slot.keep = (band, r, u, v)
            slot.done.record()
        self.pending.append((slot, host, r["q"]))

    def collect(self):
        slot, host, q = self.pending.popleft()
        slot.done.synchronize()
        s = host["state"].numpy()
        from . import _lib
        return dict(cost=[host["cost0"].numpy(), host["cost1"].numpy()], structure=[host["S0"].numpy(), host["S1"].numpy()],
                    mu=[float(s[_lib.STATE_MU]), float(s[_lib.STATE_MU + 1])], B=[float(s[_lib.STATE_B]), float(s[_lib.STATE_B + 1])],
                    epipoles=q, labels=s[_lib.STATE_LABELS:_lib.STATE_LABELS + N_RANGES].copy(), u=host["u"].numpy(),
                    v=host["v"].numpy())


def run_triplets(triplets, depth=2, **kw):
    """Generator over the results of ``triplets`` (an iterable of dicts with images, flow, rigidity,
    homographies, epipoles), ``depth`` of them in flight."""
    ts = TripletStream(depth=depth, **kw)
    for t in triplets:
        if len(ts.pending) == len(ts.slots):
            yield ts.collect()
        ts.submit(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"])
    while ts.pending:
        yield ts.collect()
