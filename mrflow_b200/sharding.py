"""Row-band sharding of one triplet over the GPUs of a box.

One process per GPU (``torch.distributed``, NCCL).  Rank r owns rows ``[y0_r, y1_r)`` of the reference
frame for all labels; per-pixel work is independent given a handful of whole-frame scalars, so the
only exchange steps are

* tiny all-reduces: the sum behind ``mu`` (pipeline/compute_structure.py:84), the counts of the
  gates (:164, :183), the label range min/max (pipeline/build_cost_volume.py:190-191);
* exact distributed medians (compute_structure.py:190, :221): an 8-pass radix select whose 256-bin
  histogram (2 KB) is all-reduced per pass -- ``DistSelect`` below, a host-side protocol around two
  per-shard device steps, so the protocol itself is testable on CPU with gloo;
* the gather of the label-major cost-volume slices to the rank that runs TRW-S: the root exports a
  whole-frame ``[L][h][w]`` buffer as peer memory and every rank's cost-volume kernel stores its
  rows straight into it over NVLink (``PeerVolume``) -- the transfer is the kernel's own output
  stream, tile by tile, not a separate collective.

Neighbour frames are uploaded per band with a halo of ``required_halo`` rows; the kernel reports
any in-frame sample whose taps fall outside it (``halo_miss``) and the band is then redone with a
larger halo.
"""
import ctypes as C
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, device, planeparallax as pp
from ._lib import check


# ------------------------------------------------------------------------------------------ plan
def band_plan(h, world, align=8):
    """Rows [y0, y1) of every rank: equal bands, boundaries aligned to the kernels' 8-row tiles."""
    units = (h + align - 1) // align
    base, extra = divmod(units, world)
    out, y = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * align
        out.append((min(y, h), min(y + n, h)))
        y += n
    return out


def required_halo(rows, w, frame_h, H, q, mu, B, labels, label_scale=1.0, stride=4):
    """Neighbour rows a band needs beyond its own: the extreme vertical displacement of the warp
    (optimize_structure_single_scale.py:210-235 + interpolate_through_H.py:31-35) over the label range
    on a ``stride``-subsampled grid of the band, plus the bicubic taps (-1..+2) and a margin for the
    subsampling.  Returns (rows above, rows below).  Host-side planning only (fp64 numpy model; the
    kernel's halo_miss counter is the authority)."""
    y0, y1 = rows
    ys = np.unique(np.r_[np.arange(y0, y1, stride), y1 - 1]).astype(np.float64)
    xs = np.unique(np.r_[np.arange(0, w, stride), w - 1]).astype(np.float64)
    x, y = np.meshgrid(xs, ys)
    vx, vy = q[0] - x, q[1] - y
    d = np.sqrt(vx ** 2 + vy ** 2)
    nrm = np.maximum(0.5, d)
    n = d / mu
    H = np.asarray(H, dtype=np.float64)
    lo, hi = 0.0, 0.0
    with np.errstate(all="ignore"):
        a = np.asarray(labels, dtype=np.float64) * label_scale
        c = (a * B) / (B * a / mu - 1.0)
        c = c[np.isfinite(c)]
        for cc in ((c.min(), c.max()) if c.size else ()):
            r = cc * n
            xn, yn = x + r * vx / nrm, y + r * vy / nrm
            den = xn * H[2, 0] + yn * H[2, 1] + H[2, 2]
            X = (xn * H[0, 0] + yn * H[0, 1] + H[0, 2]) / den
            Y = (xn * H[1, 0] + yn * H[1, 1] + H[1, 2]) / den
            ok = (X >= 0) & (X <= w - 1) & (Y >= 0) & (Y <= frame_h - 1)
            if ok.any():
                lo = min(lo, float((Y - y0)[ok].min()))
                hi = max(hi, float((Y - (y1 - 1))[ok].max()))
    margin = 3 + stride
    return int(np.ceil(-lo)) + margin, int(np.ceil(hi)) + margin


# ------------------------------------------------------------------------------------------ select
def order_key(v):
    """numpy twin of the kernels' order-preserving key (csrc/select.cu)."""
    v = np.asarray(v, dtype=np.float64)
    b = v.view(np.uint64)
    k = np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))
    return np.where(np.isnan(v), np.uint64(0xFFFFFFFFFFFFFFFF), k)


def key_value(k):
    k = int(k)
    if k == 0xFFFFFFFFFFFFFFFF:
        return float("nan")
    b = (k & 0x7FFFFFFFFFFFFFFF) if (k >> 63) else (~k & 0xFFFFFFFFFFFFFFFF)
    return float(np.array([b], dtype=np.uint64).view(np.float64)[0])


class DistSelect:
    """Exact k-th order statistic of keys sharded over ranks.

    ``hist_fn(prefix, pass) -> int64[256]`` and ``succ_fn(kth_key) -> (count_le, n_nan, min_key_gt)``
    are the per-shard steps (device kernels in production, numpy in the CPU tests);
    ``allreduce(array, op)`` sums / mins small int64 host arrays over the ranks."""

    def __init__(self, hist_fn, succ_fn, allreduce):
        self.hist_fn, self.succ_fn, self.allreduce = hist_fn, succ_fn, allreduce

    def kth(self, k):
        """Returns (kth value, (k+1)th value, total NaN count)."""
        prefix, k_rem = 0, int(k)
        for p in range(8):
            hist = self.allreduce(np.asarray(self.hist_fn(prefix, p), dtype=np.int64), "sum")
            b = 0
            while b < 255 and k_rem >= int(hist[b]):
                k_rem -= int(hist[b])
                b += 1
            prefix |= b << (56 - 8 * p)
        cnt_le, n_nan, min_gt = self.succ_fn(prefix)
        tot = self.allreduce(np.array([cnt_le, n_nan], dtype=np.int64), "sum")
        # min over ranks of an unsigned key: compare as (high, low) signed halves
        hi_lo = np.array([(min_gt >> 32) & 0xFFFFFFFF, min_gt & 0xFFFFFFFF], dtype=np.int64)
        hi = self.allreduce(hi_lo[:1].copy(), "min")
        lo = self.allreduce(np.array([hi_lo[1] if hi_lo[0] == hi[0] else 0xFFFFFFFF], dtype=np.int64), "min")
        nxt = (int(hi[0]) << 32) | int(lo[0])
        kv = key_value(prefix)
        return kv, (kv if int(tot[0]) >= k + 2 else key_value(nxt)), int(tot[1])

    def median(self, n_total):
        """numpy.median over all shards (mean of the two middles for even n, NaN if any NaN)."""
        if n_total == 0:
            return float("nan")
        a, b, n_nan = self.kth((n_total - 1) // 2)
        if n_nan > 0:
            return float("nan")
        return a if n_total % 2 == 1 else float(np.mean(np.array([a, b])))


class BandStats:
    """``planeparallax`` statistics over row bands: the ``LocalStats`` interface with collectives."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.backend = dist.get_backend(group)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")

    # -- small host arrays over the ranks
    def _allreduce(self, arr, op):
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op],
                        group=self.group)
        return t.cpu().numpy()

    def total(self, value):
        return int(self._allreduce(np.array([int(value)], dtype=np.int64), "sum")[0])

    def device_sum(self, t):
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def minmax(self, mm):
        a = mm.cpu().numpy()
        nan = self._allreduce(np.array([int(np.isnan(a).any())], dtype=np.int64), "max")[0]
        lo = self._allreduce(np.array([a[0]]), "min")[0]
        hi = self._allreduce(np.array([a[1]]), "max")[0]
        return (float("nan"), float("nan")) if nan else (float(lo), float(hi))

    def median(self, keys, n):
        L = _lib.lib()
        hist = torch.empty(256, dtype=torch.int64, device=keys.device)
        out3 = torch.empty(3, dtype=torch.int64, device=keys.device)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def hist_fn(prefix, p):
            check(L.mrf_select_hist_prefix_f64(C.c_void_p(keys.data_ptr()), int(n), C.c_ulonglong(prefix), p,
                                               C.c_void_p(hist.data_ptr()), stream), "mrf_select_hist_prefix_f64")
            return hist.cpu().numpy()

        def succ_fn(kth):
            check(L.mrf_select_successor_key_f64(C.c_void_p(keys.data_ptr()), int(n), C.c_ulonglong(kth),
                                                 C.c_void_p(out3.data_ptr()), stream), "mrf_select_successor_key_f64")
            o = out3.cpu().numpy().view(np.uint64)
            return int(o[0]), int(o[1]), int(o[2])

        return DistSelect(hist_fn, succ_fn, self._allreduce).median(self.total(n))

    def epipole(self, fn, q, band):
        """correct_epipole runs on the rank whose band contains the 21x21 window; the result is
        broadcast.  A window straddling two bands is left uncorrected on all ranks (logged)."""
        h, w = band.frame_h, band.w
        ymin, ymax = int(q[1]) - 10, int(q[1]) + 10
        xmin, xmax = int(q[0]) - 10, int(q[0]) + 10
        inside = not (xmin < 0 or xmax >= w or ymin < 0 or ymax >= h)
        if not inside:
            # compute_structure.py:476-477: the window leaves the image, the epipole stays -- every rank knows that
            # from (q, w, h) alone, so no collective and no host synchronisation
            return np.asarray(q, dtype=np.float64)
        mine = ymin >= band.y0 and ymax < band.y0 + band.rows
        res = np.zeros(3)
        if mine:
            qn = fn()
            res = np.array([1.0, qn[0], qn[1]])
        # exact transfer of the two doubles: exactly one rank contributes non-zero values
        tot = self._allreduce(res, "sum")
        if tot[0] == 0:
            if inside and self.rank == 0:
                print("(WW) epipole window straddles two row bands: epipole not corrected")
            return np.asarray(q, dtype=np.float64)
        return tot[1:3].copy()


# ------------------------------------------------------------------------------------------ gather
class PeerVolume:
    """A whole-frame ``[L][h][w]`` fp32 cost volume on the root GPU that every rank can store into.

    The root allocates it through ``mrf_peer_alloc`` (plain cudaMalloc + CUDA IPC handle); the handle is
    broadcast and opened by the other ranks, whose cost-volume kernels then write their rows with
    ``cost = base + y0*w`` and ``cost_plane_stride = h*w`` -- stores over NVLink from inside the
    compute kernel.  ``tensor()`` on the root views the buffer as a torch tensor (no copy)."""

    def __init__(self, L, h, w, root=0, group=None):
        self.L, self.h, self.w, self.root = L, h, w, root
        self.rank = dist.get_rank(group)
        self.group = group
        lib = _lib.lib()
        self.nbytes = L * h * w * 4
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        if self.rank == root:
            check(lib.mrf_peer_alloc(self.nbytes, C.byref(ptr), handle), "mrf_peer_alloc")
        obj = [bytes(handle)]
        dist.broadcast_object_list(obj, src=root, group=group)
        if self.rank != root:
            h64 = (C.c_ubyte * 64).from_buffer_copy(obj[0])
            check(lib.mrf_peer_open(h64, C.byref(ptr)), "mrf_peer_open")
        self.ptr = ptr.value

    def band_ptr(self, y0):
        return self.ptr + y0 * self.w * 4

    def tensor(self):
        assert self.rank == self.root

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.L, self.h, self.w), "typestr": "<f4", "data": (self.ptr, False),
                                      "version": 2}
        return torch.as_tensor(a, device="cuda")

    def close(self):
        lib = _lib.lib()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self.rank == self.root:
            check(lib.mrf_peer_free(C.c_void_p(self.ptr)), "mrf_peer_free")
        else:
            check(lib.mrf_peer_close(C.c_void_p(self.ptr)), "mrf_peer_close")
        self.ptr = None


# ------------------------------------------------------------------------------------------ resident bands
class PeerMailbox:
    """Every rank's small IPC-shared buffer for ``mrf_mailbox_allreduce``: the integer / min all-reduces of the
    row-band pre-pass as one kernel per rank over NVLink peer memory instead of NCCL launches (bit-identical
    results: only order-independent reductions)."""

    def __init__(self, group=None):
        lib = _lib.lib()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = lib.mrf_mailbox_bytes(self.world)
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        check(lib.mrf_peer_alloc(self.nbytes, C.byref(ptr), handle), "mrf_peer_alloc")
        self.local = ptr.value

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.local, False), "version": 2}
        torch.as_tensor(a, device="cuda").zero_()
        torch.cuda.synchronize()
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs = (C.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                self.ptrs[r] = self.local
            else:
                p = C.c_void_p()
                check(lib.mrf_peer_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p)), "mrf_peer_open")
                self.ptrs[r] = p.value
        self.calls = 0
        dist.barrier(group=group)                # every mailbox is zeroed and mapped before the first store

    def allreduce(self, words, op):
        """In-place all-reduce of a contiguous device tensor of <= 256 8-byte elements (int64 / float64)."""
        assert words.is_cuda and words.is_contiguous() and words.element_size() == 8 and words.numel() <= 256
        code = {"sum_u64": 0, "min_u64": 1, "min_f64": 2}[op]
        check(_lib.lib().mrf_mailbox_allreduce(C.c_void_p(self.local), self.ptrs, self.world, self.rank, self.calls, code,
                                               C.c_void_p(words.data_ptr()), words.numel(),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mrf_mailbox_allreduce")
        self.calls += 1

    def status(self):
        """0, or the 1-based index of the first collective that gave up waiting for a peer."""
        out = torch.empty(1, dtype=torch.int64, device="cuda")

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.nbytes // 8,), "typestr": "<i8", "data": (self.local, False), "version": 2}
        return int(torch.as_tensor(a, device="cuda")[self.nbytes // 8 - 8].cpu())

    def close(self):
        lib = _lib.lib()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for r in range(self.world):
            if r != self.rank:
                check(lib.mrf_peer_close(C.c_void_p(self.ptrs[r])), "mrf_peer_close")
        dist.barrier(group=self.group)
        check(lib.mrf_peer_free(C.c_void_p(self.local)), "mrf_peer_free")
        self.local = None


class BandBuffers:
    """Reusable HBM buffers of the banded device-resident pre-pass."""

    def __init__(self, rows, w, dev="cuda"):
        L = _lib.lib()
        F64, I32, U8 = torch.float64, torch.int32, torch.uint8
        self.par = [torch.empty((rows, w), dtype=F64, device=dev) for _ in range(2)]
        self.S = [torch.empty((rows, w), dtype=F64, device=dev) for _ in range(2)]
        self.cand = torch.empty(rows * w, dtype=F64, device=dev)
        self.count = torch.zeros(1, dtype=I32, device=dev)
        self.n_total = torch.zeros(1, dtype=I32, device=dev)
        self.sums = torch.zeros(2, dtype=F64, device=dev)
        self.state = torch.zeros(_lib.STATE_WORDS, dtype=F64, device=dev)
        self.scratch = torch.empty(L.mrf_front_scratch_bytes(), dtype=U8, device=dev)
        self.sel = torch.zeros(L.mrf_select_scratch_bytes(0), dtype=U8, device=dev)
        words = self.sel.view(torch.int64)
        self.sel_hist = words[_lib.SELECT_HIST_OFFSET // 8:_lib.SELECT_HIST_OFFSET // 8 + 256]
        self.sel_sum = words[3:5]            # NaN count, count(keys <= kth)
        self.sel_min = words[5:6]            # smallest order-key above kth (unsigned)
        self.sign = torch.tensor([-2 ** 63], dtype=torch.int64, device=dev)


def _front_params(band, homographies, q_ar, n_labels, B_fwd=1.0):
    p = _lib.FrontParams()
    p.rows, p.w, p.y0, p.frame_h = band.rows, band.w, band.y0, band.frame_h
    p.Hinv0 = device._mat(np.linalg.inv(np.asarray(homographies[0], dtype=np.float64)))
    p.Hinv1 = device._mat(np.linalg.inv(np.asarray(homographies[1], dtype=np.float64)))
    p.q0 = device._pt(q_ar[0]); p.q1 = device._pt(q_ar[1])
    p.B_fwd = float(B_fwd)
    p.n_labels = int(n_labels)
    w, h = band.w, band.frame_h
    for f in range(2):
        q = q_ar[f]
        bx0 = max(0, min(w - 1, int(q[0] - 10))); bx1 = max(0, min(w - 1, int(q[0] + 10)))
        by0 = max(0, min(h - 1, int(q[1] - 10))); by1 = max(0, min(h - 1, int(q[1] + 10)))
        p.have_box[f] = int(not (bx0 == w - 1 or bx1 == 0 or by0 == h - 1 or by1 == 0))
        for k, v in enumerate((bx0, bx1, by0, by1)):
            p.box[f][k] = v
    return p


def banded_front(band, homographies, q_ar, rigid_mask, zero_mask, buf, group=None, n_labels=pp.N_LABELS, mailbox=None):
    """The pre-pass of build_cost_volume for one row band with every whole-frame scalar kept on the device:
    kernels on this band, tiny NCCL all-reduces (2 doubles, 1 int, 8 x 2 KB histograms, 3 words, 6 doubles)
    enqueued on the same stream in between -- no host synchronisation.  Leaves mu, B, labels in
    ``buf.state`` (identical on every rank) and this band's parallax / structures in ``buf``."""
    L = _lib.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = _front_params(band, homographies, q_ar, n_labels)
    SUM, MIN, MAX = dist.ReduceOp.SUM, dist.ReduceOp.MIN, dist.ReduceOp.MAX
    fl = band.flow
    check(L.mrf_front_parallax_band(C.byref(p), P(fl[0][0]), P(fl[0][1]), P(fl[1][0]), P(fl[1][1]), P(buf.par[0]), P(buf.par[1]),
                                    P(buf.sums), P(buf.scratch), st), "mrf_front_parallax_band")
    dist.all_reduce(buf.sums, op=SUM, group=group)
    buf.state[_lib.STATE_MU:_lib.STATE_MU + 2] = buf.sums / float(band.frame_h * band.w)        # compute_structure.py:84
    check(L.mrf_front_candidates_band(C.byref(p), P(buf.par[0]), P(buf.par[1]), P(rigid_mask), P(buf.state), P(buf.cand),
                                      P(buf.count), st), "mrf_front_candidates_band")
    buf.n_total.copy_(buf.count)
    dist.all_reduce(buf.n_total, op=SUM, group=group)
    cap = band.rows * band.w
    check(L.mrf_select_begin_dev(P(buf.n_total), P(buf.sel), st), "mrf_select_begin_dev")
    for _ in range(8):
        check(L.mrf_select_hist_dev(P(buf.cand), P(buf.count), cap, P(buf.sel), st), "mrf_select_hist_dev")
        if mailbox is not None:
            mailbox.allreduce(buf.sel_hist, "sum_u64")
        else:
            dist.all_reduce(buf.sel_hist, op=SUM, group=group)
        check(L.mrf_select_pick(P(buf.sel), st), "mrf_select_pick")
    check(L.mrf_select_successor_dev(P(buf.cand), P(buf.count), cap, P(buf.sel), st), "mrf_select_successor_dev")
    if mailbox is not None:
        mailbox.allreduce(buf.sel_sum, "sum_u64")
        mailbox.allreduce(buf.sel_min, "min_u64")
    else:
        dist.all_reduce(buf.sel_sum, op=SUM, group=group)
        buf.sel_min.bitwise_xor_(buf.sign)                 # unsigned order -> signed order for MIN
        dist.all_reduce(buf.sel_min, op=MIN, group=group)
        buf.sel_min.bitwise_xor_(buf.sign)
    check(L.mrf_select_median_state(P(buf.sel), C.c_void_p(buf.state.data_ptr() + 8 * _lib.STATE_B), st),
          "mrf_select_median_state")                       # B_b, compute_structure.py:221
    check(L.mrf_front_structures_band(C.byref(p), P(buf.par[0]), P(buf.par[1]), P(zero_mask), P(buf.S[0]), P(buf.S[1]),
                                      P(buf.state), P(buf.scratch), st), "mrf_front_structures_band")
    # one MIN all-reduce for (min, -max, -nanflag) of both frames instead of three collectives (negation is exact)
    mm = buf.state[_lib.STATE_MINMAX:_lib.STATE_MINMAX + 4]
    flags = buf.state[_lib.STATE_NANFLAG:_lib.STATE_NANFLAG + 2]
    packed = torch.cat((mm[0::2], -mm[1::2], -flags))
    if mailbox is not None:
        mailbox.allreduce(packed, "min_f64")
    else:
        dist.all_reduce(packed, op=MIN, group=group)
    mm[0::2] = packed[0:2]; mm[1::2] = -packed[2:4]; flags.copy_(-packed[4:6])
    check(L.mrf_label_range_dev(P(buf.state), int(n_labels), st), "mrf_label_range_dev")
    return buf


def banded_pass_resident(band, homographies, epipoles, rigid_mask, stats, buf=None, rho="lorentzian", sigma=1.0,
                         interp="cubic", range_mask="rigid_only", variant="staged", out_ptr=(None, None), plane_stride=0,
                         want_cost=True, halo_miss=None, events=None, masks=None, channels=None, mailbox=None):
    """Row-band cost volumes with the device-resident pre-pass.  Returns dict(cost, mins, argmins, q, buffers)."""
    q_new = [stats.epipole(lambda f=f: _band_epipole(band, homographies, epipoles, rigid_mask, f), np.asarray(epipoles[f], np.float64), band)
             for f in range(2)]
    buf = buf or BandBuffers(band.rows, band.w, band.device)
    if masks is None:
        nonrigid = (rigid_mask == 0).to(torch.uint8)
        masks = (nonrigid, rigid_mask if range_mask == "as_written" else nonrigid)
    nonrigid, zero_mask = masks
    cur = torch.cuda.current_stream()
    ch_done = None
    if channels is None:
        # the channel stacks do not depend on the pre-pass: build them on a side stream while the pre-pass and
        # its collectives run
        side = _side_stream(band.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            channels = pp.channels(band)
            ch_done = torch.cuda.Event()
            ch_done.record(side)
        for tns in (channels[0], *channels[1]):
            tns.record_stream(cur)
    banded_front(band, homographies, q_new, rigid_mask, zero_mask, buf, group=stats.group, mailbox=mailbox)
    if ch_done is not None:
        cur.wait_event(ch_done)
    ref64, tex = channels
    res = pp.cost_volumes(band, ref64, tex, None, homographies, q_new, None, None, nonrigid, rho=rho, rho_param=sigma,
                          interp=interp, variant=variant, want_cost=want_cost, halo_miss=halo_miss, events=events,
                          out_ptr=out_ptr, plane_stride=plane_stride, state=buf.state)
    return dict(cost=[res[0][0], res[1][0]], mins=[res[0][1], res[1][1]], argmins=[res[0][2], res[1][2]], q=q_new, buffers=buf)


_SIDE = {}


def _side_stream(dev):
    key = str(dev)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def _band_epipole(band, homographies, epipoles, rigid_mask, f):
    q = np.asarray(epipoles[f], dtype=np.float64)
    Hinv = np.linalg.inv(np.asarray(homographies[f], dtype=np.float64))
    ur, vr, _, _ = device.flow_to_parallax(band.flow[f][0], band.flow[f][1], Hinv, q, y0=band.y0, want_parallax=False,
                                           want_dists=False)
    return pp.correct_epipole(ur, vr, q, rigid_mask, band.y0, band.frame_h)


# ------------------------------------------------------------------------------------------ driver
def banded_cost_volume(t, stats, rank, world, rho="lorentzian", sigma=1.0, interp="cubic", range_mask="rigid_only",
                       variant="staged", peer=(None, None), keep_local=True, halo=None, timers=None, resident=True,
                       mailbox=None):
    """build_cost_volume for this rank's band of the host triplet ``t`` (dict as synth.make_triplet
    returns).  ``resident=True`` keeps the whole-frame scalars on the device (``banded_pass_resident``);
    ``False`` drives the statistics through the host (``BandStats`` / ``DistSelect``, the protocol the gloo
    tests cover) -- same kernels, bit-identical results.  ``mailbox`` (a ``PeerMailbox``) routes the integer / min
    all-reduces of the resident pre-pass through NVLink peer memory instead of NCCL.
    Returns dict(cost, S, mu, B, q, labels, mins, argmins, rows, halo)."""
    from .pipeline.build_cost_volume import cost_volume_pass
    h, w = t["shape"]
    y0, y1 = band_plan(h, world)[rank]
    rig = np.asarray(t["rigidity"])
    halo_rows = 32 if halo is None else int(halo)
    while True:
        band = pp.Band.from_host(t["images"], t["flow"], rig, None, rows=(y0, y1), halo=halo_rows)
        miss = torch.zeros(1, dtype=torch.int32, device="cuda")
        outs = tuple((p.band_ptr(y0) if p is not None else None) for p in peer)
        stride = h * w if outs[0] is not None else 0
        if resident:
            mask = torch.from_numpy((rig[y0:y1] > 0).astype(np.uint8)).cuda()
            r = banded_pass_resident(band, t["homographies"], t["epipoles"], mask, stats, rho=rho, sigma=sigma, interp=interp,
                                     range_mask=range_mask, variant=variant, out_ptr=outs, plane_stride=stride,
                                     want_cost=keep_local, halo_miss=miss, events=timers, mailbox=mailbox)
            mu, B, labels = pp.state_to_host(r["buffers"].state)
            res = (r["cost"], r["buffers"].S, mu, B, r["q"], labels, r["mins"], r["argmins"])
        else:
            res = cost_volume_pass(band, t["homographies"], t["epipoles"], rig[y0:y1] > 0, stats, rho=rho, sigma=sigma,
                                   interp=interp, range_mask=range_mask, variant=variant, halo_miss=miss,
                                   want_cost=keep_local, out_ptr=outs, plane_stride=stride, events=timers)
        n_miss = stats.total(int(miss.cpu()[0]))
        if n_miss == 0:
            break
        halo_rows *= 2                      # a sample reached beyond the halo somewhere: widen everywhere
        if halo_rows > 4 * h:
            raise RuntimeError("halo search did not converge")
    cost, S, mu, B, q, labels, mins, argmins = res
    return dict(cost=cost, S=S, mu=mu, B=B, q=q, labels=labels, mins=mins, argmins=argmins, rows=(y0, y1), halo=halo_rows)


def bench_bands(args, hw, steps, warm, metric, unit, workload, ClockSampler):
    """bench.py --mode bands: ONE frame sharded by row bands over all ranks (strong scaling)."""
    from . import synth
    rank, world = dist.get_rank(), dist.get_world_size()
    h, w = hw
    L = pp.N_LABELS
    t = synth.make_triplet(hw, seed=1001)            # every rank builds the same host triplet
    stats = BandStats()
    from .pipeline.build_cost_volume import cost_volume_pass
    y0, y1 = band_plan(h, world)[rank]
    rig = np.asarray(t["rigidity"])
    # pick the halo once (planning is outside the timed region, as uploading the inputs is)
    r0 = banded_cost_volume(t, stats, rank, world, interp=args.interp, variant=args.variant, keep_local=False)
    halo_rows = r0["halo"]
    band = pp.Band.from_host(t["images"], t["flow"], rig, None, rows=(y0, y1), halo=halo_rows)
    rigid_mask = torch.from_numpy((rig[y0:y1] > 0).astype(np.uint8)).cuda()
    peers = (PeerVolume(L, h, w), PeerVolume(L, h, w))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    kernel_events = []

    nonrigid = (rigid_mask == 0).to(torch.uint8)
    # measured on 8 GPUs / 4K (profiles/r1g_scale8_bands4k*.json): the peer-memory all-reduce kernel is no faster than
    # NCCL's low-latency protocol for these 2 KB messages (2.57 ms vs 2.37 ms per step), so NCCL stays the default
    mailbox = PeerMailbox() if getattr(args, "band_allreduce", "nccl") == "peer" else None
    bbuf = BandBuffers(y1 - y0, w)
    miss = torch.zeros(1, dtype=torch.int32, device="cuda")
    cost_local = (torch.empty((L, y1 - y0, w), dtype=torch.float32, device="cuda"),
                  torch.empty((L, y1 - y0, w), dtype=torch.float32, device="cuda"))

    def step(events=None, gather=True):
        outs = tuple(p.band_ptr(y0) for p in peers) if gather else tuple(c.data_ptr() for c in cost_local)
        return banded_pass_resident(band, t["homographies"], t["epipoles"], rigid_mask, stats, buf=bbuf, interp=args.interp,
                                    range_mask="rigid_only", variant=args.variant, out_ptr=outs,
                                    plane_stride=(h * w if gather else 0), halo_miss=miss, events=events,
                                    masks=(nonrigid, nonrigid), mailbox=mailbox)

    def barrier():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    results = {}
    for mode, gather in (("compute_only", False), ("compute_and_gather", True)):
        for _ in range(warm):
            step(gather=gather)
        barrier()
        device.reset_launch_count()
        tot = 0.0
        with ClockSampler(torch.cuda.current_device()) as clocks:
            for _ in range(steps):
                flush.fill_(1)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step(kernel_events if gather else None, gather=gather)
                e1.record(); e1.synchronize()
                tot += e0.elapsed_time(e1)
        launches = device.launch_count()
        tt = torch.tensor([tot], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        results[mode] = (float(tt.cpu()[0]), launches, clocks.summary())
    units = 2 * h * w * L
    tot_ms, launches, clk = results["compute_and_gather"]
    kms = [a.elapsed_time(b) for a, b in kernel_events]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rows = y1 - y0
    bytes_launch = rows * w * (L * 4 + 4 + 12 + 2 + 6)
    ach = bytes_launch / (float(np.mean(kms)) * 1e-3) / 1e9
    barrier()
    n_miss = stats.total(int(miss.cpu()[0]))
    mbox_status = stats.total(int(mailbox.status() != 0)) if mailbox is not None else 0
    for p in peers:
        p.close()
    if mailbox is not None:
        mailbox.close()
    if rank == 0:
        assert n_miss == 0, "a timed step sampled outside its halo"
        assert mbox_status == 0, "a peer-memory all-reduce gave up waiting for a rank"
        line = {"metric": metric, "value": units * steps / (tot_ms * 1e-3), "unit": unit, "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64 geometry / f32 sampling", "data": "synthetic",
                "config": {"workload": workload, "labels": L, "frames": 2, "interp": args.interp, "variant": args.variant,
                           "range_mask": "rigid_only", "parallelism": "row bands x%d, halo %d rows, device-resident statistics "
                           "(%s), peer-store gather to rank 0" % (world, halo_rows, "NCCL all-reduces on the stream" if mailbox is None else
                                                                "histogram / min all-reduces as one kernel per rank over NVLink peer memory, two NCCL sums"),
                           "l2": "flushed between timed steps"},
                "compute_only": {"value": units * steps / (results["compute_only"][0] * 1e-3), "ms_per_step": results["compute_only"][0] / steps},
                "clocks": clk, "e2e": None, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "kernel": "k_cost_volume (this rank's band, stores to rank 0 over NVLink)",
                             "kernel_ms": float(np.mean(kms))},
                "cpu_baseline": None}
        print(json.dumps(line))
    dist.destroy_process_group()
