"""Row-band sharding of one triplet over the GPUs of a box.

One process per GPU (``torch.distributed``, NCCL for the plumbing).  Rank r owns rows ``[y0_r, y1_r)`` of the
reference frame for all labels; per-pixel work is independent given a handful of whole-frame scalars, so the
only exchange steps are

* nothing at all for ``mu`` (pipeline/compute_structure.py:84): the sum of ||p - q|| over the whole frame depends
  only on (q, h, w), so every rank evaluates it itself in the one-GPU summation tree (``banded_mu``);
* the exact median behind B_b (compute_structure.py:221): five radix passes over 13-bit digits whose 8192-bin
  histograms are summed over the ranks, one gather of the three successor words, one gather of the six words of
  the label range (pipeline/build_cost_volume.py:190-191) -- in ``banded_front`` either as seven small NCCL
  collectives enqueued on the stream, or, with a ``PeerSync``, with NO collective launch: the kernels exchange
  over NVLink peer memory themselves (csrc/peersync.cuh).  ``DistSelect13`` is the host twin of that protocol and
  ``DistSelect`` / ``BandStats`` the host-driven 8-bit protocol, both testable on CPU with gloo;
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
import sys

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


class DistSelect13:
    """Host twin of the DEVICE protocol the resident row-band pre-pass runs (csrc/select13.cuh,
    ``banded_front``): five radix passes over digits of 12+13+13+13+13 bits, one 8192-bin histogram
    all-reduce per pass (n = the total of the first one, no separate count exchange), then one all-gather of
    the three successor words (NaN count, count(keys <= kth), inverted smallest key above kth).  The per-shard
    steps are callables so the protocol is testable on CPU with gloo; the GPU tests cover the kernels.

    ``hist_fn(prefix, pass) -> int64[8192]``, ``succ_fn(kth_key) -> (n_nan, count_le, inv_min_gt)``,
    ``allreduce(array, 'sum')``, ``allgather(int-list) -> list of int-lists (one per rank)``."""
    SHIFTS = (52, 39, 26, 13, 0)
    BINS = 8192

    def __init__(self, hist_fn, succ_fn, allreduce, allgather):
        self.hist_fn, self.succ_fn, self.allreduce, self.allgather = hist_fn, succ_fn, allreduce, allgather

    def median(self):
        prefix, k_rem, n = 0, 0, 0
        for p, shift in enumerate(self.SHIFTS):
            hist = self.allreduce(np.asarray(self.hist_fn(prefix, p), dtype=np.int64), "sum")
            if p == 0:
                n = int(hist.sum())
                k_rem = (n - 1) // 2 if n else 0
            b = 0
            while b < self.BINS - 1 and k_rem >= int(hist[b]):
                k_rem -= int(hist[b])
                b += 1
            prefix = (prefix | (b << shift)) & 0xFFFFFFFFFFFFFFFF
        rows = self.allgather([int(v) for v in self.succ_fn(prefix)])
        n_nan = sum(r[0] for r in rows); cnt_le = sum(r[1] for r in rows); inv_min = max(r[2] for r in rows)
        if n == 0 or n_nan > 0:
            return float("nan")
        a = key_value(prefix)
        k = (n - 1) // 2
        b = a if cnt_le >= k + 2 else key_value(~inv_min & 0xFFFFFFFFFFFFFFFF)
        return a if n % 2 == 1 else float(np.mean(np.array([a, b])))


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

    def _allgather(self, values):
        """small unsigned 64-bit integers, one list per rank"""
        mine = torch.from_numpy(np.array([v & 0xFFFFFFFFFFFFFFFF for v in values], dtype=np.uint64).view(np.int64)).to(self.dev)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return [[int(x) for x in o.cpu().numpy().view(np.uint64)] for o in out]

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

    def window(self, win):
        """The 21x21 epipole window (planeparallax.epipole_window) summed over the bands: every rank holds
        the rows it owns and zeros elsewhere, so the sum is exact and every rank then solves the same
        441-pixel problem (compute_structure.py:453-511) -- no broadcast, and a window that straddles a band
        boundary is handled like any other."""
        return self._allreduce(np.ascontiguousarray(win), "sum")


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
    """Every rank's small IPC-shared buffer for ``mrf_mailbox_allreduce``: integer / min all-reduces of up to 256
    words as one kernel per rank over NVLink peer memory (bit-identical to NCCL: only order-independent
    reductions).  A building block: the row-band pre-pass itself uses NCCL (its histograms are 32 KB).  A rank
    that never shows up turns the result into all-ones after 20 s and sets the status word: call ``check()``
    (a collective) before trusting results."""

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

    def check(self):
        """Raise on every rank if any rank's collective gave up waiting for a peer (results are poisoned then)."""
        bad = torch.tensor([int(self.status() != 0)], dtype=torch.int64, device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if int(bad.cpu()[0]):
            raise RuntimeError("a peer-memory all-reduce timed out waiting for a rank; its results are poisoned")

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


class PeerSync:
    """The peer-memory context of the row-band pre-pass (csrc/peersync.cuh): every rank's IPC-shared block of
    flags, gather slots and double-buffered histograms, mapped by all ranks.  With it the pre-pass runs WITHOUT
    collective launches: its kernels reduce into every rank's histogram over NVLink, publish a flag, and the
    next kernel polls its own flags.  Create once per process group (a collective), pass to ``banded_front`` /
    ``banded_pass_resident``; every rank must then run the same sequence of steps.  ``check()`` (a collective)
    raises if a rank timed out waiting for a peer."""

    def __init__(self, group=None):
        lib = _lib.lib()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerSync supports up to 8 ranks (one NVSwitch box)")
        self.nbytes = lib.mrf_peer_sync_bytes()
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        check(lib.mrf_peer_alloc(self.nbytes, C.byref(ptr), handle), "mrf_peer_alloc")
        self.local = ptr.value
        self._view().zero_()
        torch.cuda.synchronize()
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.arg = _lib.PeerSyncArg()
        self.arg.world, self.arg.rank = self.world, self.rank
        for r in range(self.world):
            if r == self.rank:
                self.arg.ctx[r] = self.local
            else:
                p = C.c_void_p()
                check(lib.mrf_peer_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p)), "mrf_peer_open")
                self.arg.ctx[r] = p.value
        dist.barrier(group=group)                # every context is zeroed and mapped before the first step

    def _view(self):
        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.nbytes // 8,), "typestr": "<i8", "data": (self.local, False), "version": 2}
        return torch.as_tensor(a, device="cuda")

    def ref(self):
        return C.byref(self.arg)

    def begin(self):
        """start of a step on the current stream: advances the device-side epoch"""
        check(_lib.lib().mrf_peer_sync_begin(C.byref(self.arg), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "mrf_peer_sync_begin")

    def status(self):
        """(epoch, error) of this rank; error != 0: 1 + the phase whose wait timed out"""
        w = self._view()[:2].cpu()
        return int(w[0]), int(w[1])

    def check(self):
        bad = torch.tensor([int(self.status()[1] != 0)], dtype=torch.int64, device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if int(bad.cpu()[0]):
            raise RuntimeError("a rank timed out waiting for a peer in the row-band pre-pass; results of that step are invalid")

    def close(self):
        lib = _lib.lib()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for r in range(self.world):
            if r != self.rank:
                check(lib.mrf_peer_close(C.c_void_p(self.arg.ctx[r])), "mrf_peer_close")
        dist.barrier(group=self.group)
        check(lib.mrf_peer_free(C.c_void_p(self.local)), "mrf_peer_free")
        self.local = None


class BandBuffers:
    """Reusable HBM buffers of the banded device-resident pre-pass."""

    def __init__(self, rows, w, dev="cuda", world=1):
        L = _lib.lib()
        F64, I32, U8, I64 = torch.float64, torch.int32, torch.uint8, torch.int64
        self.par = [torch.empty((rows, w), dtype=F64, device=dev) for _ in range(2)]
        self.S = [torch.empty((rows, w), dtype=F64, device=dev) for _ in range(2)]
        self.cand = torch.empty(rows * w, dtype=F64, device=dev)
        self.count = torch.zeros(1, dtype=I32, device=dev)
        self.state = torch.zeros(_lib.STATE_WORDS, dtype=F64, device=dev)
        self.scratch = torch.empty(L.mrf_front_scratch_bytes(), dtype=U8, device=dev)
        self.mu_scratch = torch.empty(L.mrf_front_scratch_bytes(), dtype=U8, device=dev)
        self.sel = torch.zeros(L.mrf_select_scratch_bytes(0), dtype=U8, device=dev)
        w32 = self.sel.view(I32)
        o, nb = _lib.SELECT13_HIST_OFFSET // 4, _lib.SELECT13_BINS
        self.sel_hist = [w32[o + p * nb:o + (p + 1) * nb] for p in range(_lib.SELECT13_PASSES)]   # one 32 KB histogram per pass
        self.sel_succ = self.sel.view(I64)[_lib.SELECT13_SUCC_OFFSET // 8:_lib.SELECT13_SUCC_OFFSET // 8 + 3]
        self.succ_all = torch.zeros(3 * world, dtype=I64, device=dev)
        self.mm = self.state[_lib.STATE_MINMAX:_lib.STATE_MINMAX + 6]          # min0 max0 min1 max1 nan0 nan1
        self.mm_all = torch.zeros(6 * world, dtype=F64, device=dev)
        self.mu_done = torch.cuda.Event()


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


def banded_mu(band, homographies, q_ar, buf, n_labels=pp.N_LABELS):
    """mu of both frames (compute_structure.py:84) on this rank, no collective: the sum of ||p - q|| over
    the WHOLE frame depends only on (q, h, w), and is evaluated in the one-GPU front's grid and summation
    tree -- bit-identical to a single-GPU run.  Enqueued on the current stream; records ``buf.mu_done``."""
    p = _front_params(band, homographies, q_ar, n_labels)
    check(_lib.lib().mrf_front_mu_whole(C.byref(p), C.c_void_p(buf.state.data_ptr()), C.c_void_p(buf.mu_scratch.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mrf_front_mu_whole")
    buf.mu_done.record()


def banded_front(band, homographies, q_ar, rigid_mask, zero_mask, buf, group=None, n_labels=pp.N_LABELS, mu_ready=False,
                 peer=None):
    """The pre-pass of build_cost_volume for one row band with every whole-frame scalar kept on the device and no
    host synchronisation.  The exact median (13-bit digits) needs five histogram sums over the ranks, its two
    middle order statistics one gather of 3 words, the label range one gather of 6; mu needs nothing
    (``banded_mu``).  ``peer=None``: SEVEN small NCCL collectives enqueued on the stream between the kernels.
    ``peer`` (a ``PeerSync``): NO collective launch -- the kernels exchange over NVLink peer memory themselves
    (remote reductions into every rank's histogram, flag stores, the next kernel polls).  Same bits either way.
    Leaves mu, B, labels in ``buf.state`` (identical on every rank) and this band's parallax / structures in
    ``buf``."""
    L = _lib.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = _front_params(band, homographies, q_ar, n_labels)
    world = dist.get_world_size(group)
    pr = peer.ref() if peer is not None else None
    fl = band.flow
    if peer is not None:
        peer.begin()
    check(L.mrf_front_parallax_band(C.byref(p), P(fl[0][0]), P(fl[0][1]), P(fl[1][0]), P(fl[1][1]), P(buf.par[0]), P(buf.par[1]),
                                    P(buf.scratch), st), "mrf_front_parallax_band")
    if mu_ready:
        torch.cuda.current_stream().wait_event(buf.mu_done)
    else:
        banded_mu(band, homographies, q_ar, buf, n_labels)
    check(L.mrf_front_candidates_band(C.byref(p), P(buf.par[0]), P(buf.par[1]), P(rigid_mask), P(buf.state), P(buf.cand),
                                      P(buf.count), P(buf.sel), pr, st), "mrf_front_candidates_band")
    cap = band.rows * band.w
    for ps in range(_lib.SELECT13_PASSES):
        if peer is None:
            dist.all_reduce(buf.sel_hist[ps], op=dist.ReduceOp.SUM, group=group)
        else:
            check(L.mrf_peer_push_hist(pr, P(buf.sel), ps, st), "mrf_peer_push_hist")
        if ps + 1 < _lib.SELECT13_PASSES:
            check(L.mrf_select13_hist_dev(P(buf.cand), P(buf.count), cap, ps + 1, P(buf.sel), pr, st), "mrf_select13_hist_dev")
        else:
            check(L.mrf_select13_successor_dev(P(buf.cand), P(buf.count), cap, P(buf.sel), pr, st), "mrf_select13_successor_dev")
    if peer is None:
        dist.all_gather_into_tensor(buf.succ_all, buf.sel_succ, group=group)
    # B_b = the median (compute_structure.py:221), structures, this band's range min / max
    check(L.mrf_front_structures_band(C.byref(p), P(buf.par[0]), P(buf.par[1]), P(zero_mask), P(buf.S[0]), P(buf.S[1]),
                                      P(buf.state), P(buf.sel), None if peer is not None else P(buf.succ_all), world,
                                      P(buf.scratch), pr, st), "mrf_front_structures_band")
    if peer is None:
        dist.all_gather_into_tensor(buf.mm_all, buf.mm, group=group)
    check(L.mrf_label_range_dev(P(buf.state), int(n_labels), None if peer is not None else P(buf.mm_all), world, pr, st),
          "mrf_label_range_dev")
    return buf


def banded_pass_resident(band, homographies, epipoles, rigid_mask, stats, buf=None, rho="lorentzian", sigma=1.0,
                         interp="cubic", range_mask="rigid_only", variant="staged", out_ptr=(None, None), plane_stride=0,
                         want_cost=True, halo_miss=None, events=None, masks=None, channels=None, q_new=None, peer=None):
    """Row-band cost volumes with the device-resident pre-pass.  Returns dict(cost, mins, argmins, q, buffers).
    ``q_new``: already refined epipoles (then nothing in here touches the host, so the whole step -- kernels,
    NCCL collectives, the side-stream fork -- can be captured in a CUDA graph)."""
    if q_new is None:
        q_new = pp.refine_epipoles(band, homographies, epipoles, rigid_mask, stats)
    buf = buf or BandBuffers(band.rows, band.w, band.device, world=stats.world)
    if masks is None:
        nonrigid = (rigid_mask == 0).to(torch.uint8)
        masks = (nonrigid, rigid_mask if range_mask == "as_written" else nonrigid)
    nonrigid, zero_mask = masks
    cur = torch.cuda.current_stream()
    # mu (whole-frame, no collective) and the channel stacks do not depend on the band's pre-pass: they run on a
    # side stream while the pre-pass and its collectives run
    side = _side_stream(band.device)
    side.wait_stream(cur)
    ch_done = None
    with torch.cuda.stream(side):
        banded_mu(band, homographies, q_new, buf)
        if channels is None:
            channels = pp.channels(band)
            ch_done = torch.cuda.Event()
            ch_done.record(side)
            for tns in (channels[0], *channels[1]):
                tns.record_stream(cur)
    banded_front(band, homographies, q_new, rigid_mask, zero_mask, buf, group=stats.group, mu_ready=True, peer=peer)
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


# ------------------------------------------------------------------------------------------ driver
def banded_cost_volume(t, stats, rank, world, rho="lorentzian", sigma=1.0, interp="cubic", range_mask="rigid_only",
                       variant="staged", peer=(None, None), keep_local=True, halo=None, timers=None, resident=True,
                       peer_sync=None):
    """build_cost_volume for this rank's band of the host triplet ``t`` (dict as synth.make_triplet
    returns).  ``resident=True`` keeps the whole-frame scalars on the device (``banded_pass_resident``);
    ``False`` drives the statistics through the host (``BandStats`` / ``DistSelect``, the protocol the gloo
    tests cover) -- same kernels, bit-identical results.
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
                                     want_cost=keep_local, halo_miss=miss, events=timers, peer=peer_sync)
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


def device_band(td, rows, halo):
    """``planeparallax.Band`` of rows ``(y0, y1)`` cut from a device-resident triplet (``synth.make_triplet_device``),
    with ``halo`` neighbour rows and the 2-row stencil margin of the channel construction."""
    h, w = td["shape"]
    y0, y1 = rows
    n0, n1 = max(0, y0 - halo), min(h, y1 + halo)
    m0, m1 = max(0, n0 - 2), min(h, n1 + 2)
    imgs = [I[m0:m1].contiguous() for I in td["images"]]
    fl = [[c[y0:y1].contiguous() for c in td["flow"][f]] for f in range(2)]
    return pp.Band(imgs, fl, td["rigidity"][y0:y1].contiguous(), None, y0=y0, frame_h=h, nb_y0=n0, nb_rows=n1 - n0, img_y0=m0)


def bench_bands(hw, steps, warm, interp="cubic", variant="staged", ClockSampler=None, verify=True, seed=1001, n_waves=16,
                use_graph=True):
    """ONE frame (BASELINE.json configs[3]: 4K) sharded by row bands over all ranks -- strong scaling; with one
    rank the whole frame on one GPU, which is the base of the speed-up.  The synthetic triplet is generated on
    rank 0's GPU and broadcast.  Timed per step (CUDA events, L2 flushed, max over ranks):
      compute_only        every rank keeps its band of both volumes;
      compute_and_gather  the bands are stored straight into rank 0's whole-frame volumes over NVLink
                          (PeerVolume: the gather is the kernel's own output stream);
      compute_and_d2h     every rank copies its band to pinned host memory over its own PCIe link (what a
                          host-side consumer such as TRW-S needs: N links instead of one).
    ``verify``: the gathered volumes equal a whole-frame run on rank 0 bit for bit, and mu / B / labels equal
    the one-GPU pre-pass's.  Returns the record on rank 0 (None elsewhere)."""
    from . import synth
    from .pipeline import build_cost_volume as bcv
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    h, w = hw
    L = pp.N_LABELS
    td = synth.make_triplet_device(hw, seed=seed, n_waves=n_waves)
    if world > 1:
        for tns in (*td["images"], *td["flow"][0], *td["flow"][1]):
            dist.broadcast(tns, 0)
    H_ar, q_ar = td["homographies"], td["epipoles"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step):
        for _ in range(warm):
            step()
        barrier()
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record(); e1.synchronize()
            tot += e0.elapsed_time(e1)
        tt = torch.tensor([tot], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.cpu()[0]) / steps

    units = 2 * h * w * L
    rec = {"workload": "4k_%dx%d_L%d_F2_%s" % (w, h, L, interp), "n_gpus": world, "steps": steps, "scaling": "strong",
           "unit": "pixel*labels/s"}
    if world == 1:
        band = device_band(td, (0, h), 0)
        rigid = (td["rigidity"] > 0).to(torch.uint8)
        st = bcv.ResidentStep(band, H_ar, q_ar, rigid, interp=interp, range_mask="rigid_only", variant=variant, use_graph=use_graph)
        ms = timed(lambda: st.run())
        rec.update({"parallelism": "whole frame on one GPU", "compute_only": {"ms_per_step": ms, "value": units / (ms * 1e-3)}})
        return rec

    stats = BandStats()
    ps = PeerSync()
    y0, y1 = band_plan(h, world)[rank]
    rows = y1 - y0
    halo_rows = 32
    while True:                                      # planning, outside the timed region (as uploading the inputs is)
        band = device_band(td, (y0, y1), halo_rows)
        rigid = (td["rigidity"][y0:y1] > 0).to(torch.uint8).contiguous()
        nonrigid = (rigid == 0).to(torch.uint8)
        miss = torch.zeros(1, dtype=torch.int32, device="cuda")
        bbuf = BandBuffers(rows, w, world=world)
        banded_pass_resident(band, H_ar, q_ar, rigid, stats, buf=bbuf, interp=interp, variant=variant, want_cost=False,
                             halo_miss=miss, masks=(nonrigid, nonrigid), peer=ps)
        if stats.total(int(miss.cpu()[0])) == 0:
            break
        halo_rows *= 2
        if halo_rows > 4 * h:
            raise RuntimeError("halo search did not converge")
    peers = (PeerVolume(L, h, w), PeerVolume(L, h, w))
    cost_local = tuple(torch.empty((L, rows, w), dtype=torch.float32, device="cuda") for _ in range(2))
    host = tuple(torch.empty((L, rows, w), dtype=torch.float32, pin_memory=True) for _ in range(2))
    kernel_events = []

    q_new = pp.refine_epipoles(band, H_ar, q_ar, rigid, stats)

    def eager(mode, events=None):
        gather = mode == "gather"
        outs = tuple(p.band_ptr(y0) for p in peers) if gather else tuple(c.data_ptr() for c in cost_local)
        return banded_pass_resident(band, H_ar, q_ar, rigid, stats, buf=bbuf, interp=interp, range_mask="rigid_only",
                                    variant=variant, out_ptr=outs, plane_stride=(h * w if gather else 0), halo_miss=miss,
                                    events=events, masks=(nonrigid, nonrigid), q_new=q_new,
                                    peer=None if mode == "local_nccl" else ps)

    # The step is ~25 short launches and 7 collectives: enqueued from Python it is bound by the host, not the GPUs.
    # Capture it once per output mode (NCCL collectives are capturable) and replay with one launch.
    graphs = {}
    if use_graph:
        try:
            for mode in ("local", "gather", "local_nccl"):
                for _ in range(2):
                    eager(mode)
                barrier()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eager(mode)
                graphs[mode] = g
                barrier()
        except Exception as e:                               # noqa: BLE001
            if rank == 0:
                print("(WW) CUDA graph capture of the banded step failed, running eagerly: %s" % e, file=sys.stderr)
            graphs = {}
            torch.cuda.synchronize()

    def step(mode, events=None):
        key = mode if mode in ("gather", "local_nccl") else "local"
        if key in graphs and events is None:
            graphs[key].replay()
        else:
            eager(key, events)
        if mode == "d2h":
            for f in range(2):
                host[f].copy_(cost_local[f], non_blocking=True)

    res = {}
    res["compute_only"] = timed(lambda: step("local"))
    res["compute_and_gather"] = timed(lambda: step("gather"))
    res["compute_and_d2h"] = timed(lambda: step("d2h"))
    res["compute_only_nccl_collectives"] = timed(lambda: step("local_nccl"))
    for _ in range(3):                                       # the kernel alone: eager launches with events round it
        flush.fill_(1)
        eager("local", kernel_events)
    torch.cuda.synchronize()
    if not graphs:
        res_eager = None
    else:
        res_eager = timed(lambda: eager("local"))
    barrier()
    n_miss = stats.total(int(miss.cpu()[0]))
    kms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    for k, ms in res.items():
        rec[k] = {"ms_per_step": ms, "value": units / (ms * 1e-3)}
    if res_eager is not None:
        rec["compute_only_eager_launches"] = {"ms_per_step": res_eager, "value": units / (res_eager * 1e-3)}
    ps.check()
    rec.update({"parallelism": "row bands x%d, device-resident statistics exchanged over NVLink peer memory from inside the "
                               "pre-pass kernels (5 histogram reductions, 2 small gathers; no collective launch; none at all "
                               "for mu); compute_only_nccl_collectives = the same step with 7 NCCL collectives instead; %s" % (
                                   world, "the step replayed as one CUDA graph" if graphs else "eager launches"),
                "peer_sync_epochs": ps.status()[0],
                "halo_rows": halo_rows, "n_miss": int(n_miss), "kernel_ms_per_launch": kms, "launches_per_step": 1,
                "band_rows": rows, "d2h_bytes_per_rank": int(2 * L * rows * w * 4)})
    if verify:
        step("gather")
        barrier()
        ok = None
        if rank == 0:
            whole = device_band(td, (0, h), 0)
            rig_w = (td["rigidity"] > 0).to(torch.uint8)
            one = bcv.cost_volume_pass_resident(whole, H_ar, q_ar, rig_w, interp=interp, range_mask="rigid_only", variant=variant)
            lab = slice(_lib.STATE_LABELS, _lib.STATE_LABELS + L)
            ok = bool(torch.equal(one["buffers"].state[:4], bbuf.state[:4]) and torch.equal(one["buffers"].state[lab], bbuf.state[lab]))
            for f in range(2):
                ok = ok and bool(torch.equal(peers[f].tensor(), one["cost"][f]))
            del one
        rec["bands_bit_exact"] = ok
    barrier()
    for p in peers:
        p.close()
    ps.close()
    if n_miss:
        raise RuntimeError("a timed step sampled outside its halo")
    return rec if rank == 0 else None
