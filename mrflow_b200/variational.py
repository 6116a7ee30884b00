"""Variational structure refinement, device-resident.

structure_refinement/optimize_structure_single_scale.py:310-652: per outer iteration two linearised data
terms, image-adaptive robust first / second order smoothness, an optional consistency term, one linear
solve, clipping and a 7x7 median.  The reference assembles scipy.sparse matrices and calls
scipy.sparse.linalg.minres (:86, :593); here the operator is applied matrix-free by one stencil kernel
(``mrf_var_apply``) and MINRES runs on the device, scalar recurrences and stopping tests included
(``mrf_minres_iterate``: one persistent cooperative kernel, two grid barriers per iteration); the host reads
the solver state once per solve.

Everything numeric is fp64 like the reference; the summation order inside the operator and the dot
products differs from scipy's CSR products, so results agree to solver tolerance (tests state 1e-6 on the
structure), not bit for bit.
"""
import numpy as np
import torch

from . import _lib, device
from ._lib import check, dvec
from .device import F32, F64, U8, _chk, _p, _stream

EPS_DIAG = 0.01          # :591 -- the constant added to the diagonal of the system


def gray32(rgb):
    """cv2.cvtColor(I.astype('float32'), COLOR_RGB2GRAY) (:381-383) as an fp32 [h,w] device tensor."""
    h, w = rgb.shape[:2]
    out = torch.empty((h, w), dtype=F32, device=rgb.device)
    check(_lib.lib().mrf_gray32(_p(rgb), int(rgb.dtype == F64), h * w, _p(out), _stream()), "mrf_gray32")
    return out


def edge_weights_local(gray, rigidity_eroded=None, norm_radius=45, min_weight=0.0):
    """utils/edge_weights.py:38-55 on the fp32 gray image, then :388-392 (zero outside the eroded rigid
    region, floor at min_weight).  Returns (weights_x, weights_y) fp32 [h,w]."""
    _chk(gray, F32, "gray")
    h, w = gray.shape
    dev = gray.device
    wx = torch.empty((h, w), dtype=F32, device=dev); wy = torch.empty_like(wx)
    ta = torch.empty_like(wx); tb = torch.empty_like(wx)
    ra = torch.empty((h, w), dtype=F64, device=dev); rb = torch.empty_like(ra)
    check(_lib.lib().mrf_edge_weights_local(_p(gray), _p(rigidity_eroded), h, w, int(norm_radius), float(min_weight),
                                            _p(wx), _p(wy), _p(ta), _p(tb), _p(ra), _p(rb), _stream()),
          "mrf_edge_weights_local")
    return wx, wy


class _Scalar:
    """One device double + pinned mirror: the dot products MINRES needs on the host."""

    def __init__(self, dev, n=1):
        self.d = torch.empty(n, dtype=F64, device=dev)
        self.h = torch.empty(n, dtype=F64, pin_memory=True)

    def get(self):
        self.h.copy_(self.d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h.numpy()


class Operator:
    """A = Z (D + l1 G'P1G + l2 G2'P2G2 + lc Pc) Z + eps I  (:580-595) over device planes."""

    def __init__(self, h, w, rigid, D, p1x, p1y, p2x, p2y, pc, l1, l2, lc, eps=EPS_DIAG):
        self.h, self.w = h, w
        self.args = (rigid, D, p1x, p1y, p2x, p2y, pc)
        self.l = (float(l1), float(l2), float(lc), float(eps))

    def apply(self, x, y):
        rigid, D, p1x, p1y, p2x, p2y, pc = self.args
        check(_lib.lib().mrf_var_apply(_p(x), _p(rigid), _p(D), _p(p1x), _p(p1y), _p(p2x), _p(p2y), _p(pc), *self.l,
                                       self.h, self.w, _p(y), _stream()), "mrf_var_apply")
        return y


MINRES_BATCH = 2000      # iteration bound of one launch of the persistent solver kernel


class MinresWorkspace:
    """Vectors and state of one MINRES solve of size h*w, reusable across solves."""

    def __init__(self, h, w, dev="cuda"):
        L = _lib.lib()
        n = h * w
        self.h, self.w, self.n = h, w, n
        self.r3 = torch.empty((3, n), dtype=F64, device=dev)
        self.w2 = torch.empty((2, n), dtype=F64, device=dev)
        self.v2 = torch.empty((2, n), dtype=F64, device=dev)
        self.x = torch.empty(n, dtype=F64, device=dev)
        self.state = _Scalar(dev, _lib.MINRES_STATE_WORDS)
        self.scratch = torch.empty(L.mrf_minres_scratch_bytes(h, w), dtype=U8, device=dev)


def minres(op, b, rtol=1e-5, maxiter=None, ws=None, batch=MINRES_BATCH):
    """scipy.sparse.linalg.minres(A, b, rtol=...) with x0 = 0 and no preconditioner -- the call behind
    ``solve(A, b, 'minres', tol)`` (:86).  Same recurrences, same stopping tests, same iteration cap (5 n),
    all evaluated on the device by one persistent cooperative kernel (``mrf_minres_iterate``) that runs until
    the solve stops or ``batch`` iterations have passed; the host then reads the 256-byte state block.
    Returns (x, info dict); x is ``ws.x`` (valid until the workspace's next solve)."""
    L = _lib.lib()
    n = b.numel()
    ws = ws or MinresWorkspace(op.h, op.w, b.device)
    if maxiter is None:
        maxiter = 5 * n
    check(L.mrf_minres_init(_p(b), n, _p(ws.r3), _p(ws.w2), _p(ws.x), _p(ws.state.d), _p(ws.scratch), _stream()),
          "mrf_minres_init")
    rigid, D, p1x, p1y, p2x, p2y, pc = op.args
    done = 0
    st = ws.state.get() if maxiter <= 0 else None
    while maxiter > 0:
        k = int(min(batch, maxiter - done))
        check(L.mrf_minres_iterate(_p(rigid), _p(D), _p(p1x), _p(p1y), _p(p2x), _p(p2y), _p(pc), *op.l, op.h, op.w, _p(ws.r3),
                                   _p(ws.w2), _p(ws.v2), _p(ws.x), _p(ws.state.d), _p(ws.scratch), done, k, float(rtol),
                                   int(maxiter), _stream()), "mrf_minres_iterate")
        done += k
        st = ws.state.get()
        if st[_lib.MINRES_ISTOP] != 0 or done >= maxiter:
            break
    istop = int(st[_lib.MINRES_ISTOP])
    if istop == 7:
        # NaN / inf in the system: scipy's iterates are all-NaN from the first dot product on (and the reference
        # then zeroes the update, :600-606); same outcome here without running 5n iterations on NaNs
        ws.x.fill_(float("nan"))
    return ws.x, dict(istop=0 if istop == 100 else istop, itn=int(st[_lib.MINRES_ITN]), rnorm=float(st[_lib.MINRES_RNORM]),
                      Anorm=float(st[_lib.MINRES_ANORM]), Acond=float(st[_lib.MINRES_ACOND]), enqueued=done)


class Refiner:
    """One frame triplet at one scale: the constant device planes (channel stacks, derivative texels,
    edge weights, masks) and ``step()`` = one outer iteration of :419-650."""

    def __init__(self, images, occlusions_bwd, occlusions_fwd, rigidity, H_ar, B_ar, q_ar, mu_ar, A_bwd_reference,
                 A_fwd_reference, lambda_1storder, lambda_2ndorder, lambda_consistency, A_scale_factor=1.0,
                 min_weight=0.0, channel_weights=(0.01, 1.0, 1.0), rtol=1e-5):
        """``images`` three contiguous CUDA [h,w,3] tensors (fp64 or fp32 RGB in [0,1]); masks uint8 CUDA
        [h,w] (rigidity = 1 where rigid); structures fp64 CUDA [h,w]; the motion data as host values."""
        self.h, self.w = images[1].shape[:2]
        self.n = self.h * self.w
        self.dev = images[1].device
        self.ch = [device.prepare_channels(I, want_texels=False)[0] for I in images]      # :350-371
        self.cw = tuple(float(c) for c in channel_weights)
        self.rigid = rigidity
        self.nonrigid = (rigidity == 0).to(U8)
        self.occ_b, self.occ_f = occlusions_bwd, occlusions_fwd
        self.H, self.B, self.q, self.mu = H_ar, [float(b) for b in B_ar], q_ar, [float(m) for m in mu_ar]
        self.prep_f = device.data_term_prepare(self.ch[1], self.ch[2], H_ar[1])
        self.prep_b = device.data_term_prepare(self.ch[1], self.ch[0], H_ar[0])
        eroded = device.erode3x3(rigidity)                                               # :387
        self.wx, self.wy = edge_weights_local(gray32(images[1]), eroded, 45, min_weight)  # :381-392
        self.A_bwd_ref, self.A_fwd_ref = A_bwd_reference, A_fwd_reference
        self.l1, self.l2, self.lc = float(lambda_1storder), float(lambda_2ndorder), float(lambda_consistency)
        self.afac_sq = float(A_scale_factor) ** 2
        self.rtol = rtol
        L = _lib.lib()
        self.scratch = torch.empty(L.mrf_var_scratch_bytes(), dtype=U8, device=self.dev)
        self.energies = _Scalar(self.dev, 5)             # E_fwd, E_bwd, E_1st, E_2nd, E_cons (:564)
        self.med = torch.zeros(3, dtype=F64, device=self.dev)   # medians behind the auto-sigmas: 1st, 2nd order, consistency
        z = lambda: torch.empty((self.h, self.w), dtype=F64, device=self.dev)
        self.buf = {k: z() for k in ("arg1", "arg2", "root1", "root2", "p1x", "p1y", "p2x", "p2y", "D", "bd", "Izc", "argc",
                                     "rootc", "pc", "sA", "b", "Apd", "med")}
        self.ws = MinresWorkspace(self.h, self.w, self.dev)
        self.last = None

    def step(self, A):
        """One outer iteration on ``A`` (fp64 CUDA [h,w], updated in place).  Returns the energy (:564).  The
        only host synchronisations are the read of the solver state after the MINRES launch and of the five
        energy terms at the end; the auto-sigmas stay on the device as medians."""
        L = _lib.lib()
        h, w, n, B = self.h, self.w, self.n, self.buf
        s_bwd = self.mu[0] / self.mu[1]
        check(L.mrf_vec_scale(_p(A), n, s_bwd, _p(B["med"]), _stream()), "mrf_vec_scale")           # A * mu0/mu1 (:436)
        df, bf, Ef, _ = device.data_term(A, self.ch[1], self.ch[2], self.nonrigid, self.occ_f, self.H[1], self.B[1],
                                         self.mu[1], self.q[1], cw=self.cw, prepared=self.prep_f)   # :425-432
        db, bb, Eb, _ = device.data_term(B["med"], self.ch[1], self.ch[0], self.nonrigid, self.occ_b, self.H[0], self.B[0],
                                         self.mu[0], self.q[0], cw=self.cw, prepared=self.prep_b)   # :435-442
        check(L.mrf_var_data_merge(_p(df), _p(bf), _p(db), _p(bb), _p(self.occ_f), _p(self.occ_b), _p(A), _p(self.A_bwd_ref),
                                   _p(self.A_fwd_ref), n, s_bwd, self.mu[1] / self.mu[0], self.afac_sq, _p(B["D"]), _p(B["bd"]),
                                   _p(B["Izc"]), _p(B["argc"]), _p(B["rootc"]), _stream()), "mrf_var_data_merge")
        pc = None
        if self.lc != 0.0:
            device.median_dev(B["rootc"].view(-1), out=self.med[2:3])                               # :491-497
            check(L.mrf_var_consistency_weights(_p(B["argc"]), n, self.afac_sq, 0.0, _p(self.med[2:3]), _p(B["pc"]),
                                                _p(self.energies.d[4:5]), _p(self.scratch), _stream()),
                  "mrf_var_consistency_weights")
            pc = B["pc"]
        check(L.mrf_var_smooth_args(_p(A), _p(self.wx), _p(self.wy), h, w, self.afac_sq, _p(B["arg1"]), _p(B["arg2"]),
                                    _p(B["root1"]), _p(B["root2"]), _stream()), "mrf_var_smooth_args")
        device.median_dev(B["root1"].view(-1), out=self.med[0:1])                                   # auto-sigma medians,
        device.median_dev(B["root2"].view(-1), out=self.med[1:2])                                   # robust_functions.py:23-29
        check(L.mrf_var_smooth_weights(_p(B["arg1"]), _p(B["arg2"]), _p(self.wx), _p(self.wy), n, self.afac_sq, 0.0, 0.0,
                                       _p(self.med), _p(B["p1x"]), _p(B["p1y"]), _p(B["p2x"]), _p(B["p2y"]),
                                       _p(self.energies.d[2:4]), _p(self.scratch), _stream()), "mrf_var_smooth_weights")
        # right-hand side: (l1 A1 + l2 A2) A without the rigidity mask, then b = Z(-bd - ... ) (:581-588)
        smooth = Operator(h, w, None, None, B["p1x"], B["p1y"], B["p2x"], B["p2y"], None, self.l1, self.l2, 0.0, eps=0.0)
        smooth.apply(A, B["sA"])
        check(L.mrf_var_rhs(_p(B["bd"]), _p(B["sA"]), _p(pc), _p(B["Izc"]), _p(self.rigid), n, self.lc, _p(B["b"]), _stream()),
              "mrf_var_rhs")
        op = Operator(h, w, self.rigid, B["D"], B["p1x"], B["p1y"], B["p2x"], B["p2y"], pc, self.l1, self.l2, self.lc)
        self.energies.d[0:1].copy_(Ef); self.energies.d[1:2].copy_(Eb)
        da, info = minres(op, B["b"], rtol=self.rtol, ws=self.ws)                                   # :593
        check(L.mrf_var_clip_update(_p(da), _p(A), h, w, dvec(self.q[1]), self.mu[1], self.B[1], 1.0, _p(B["Apd"]), _stream()),
              "mrf_var_clip_update")                                                                # :600-612
        check(L.mrf_median7x7_f64(_p(B["Apd"]), h, w, _p(B["med"]), _stream()), "mrf_median7x7_f64")   # :636
        check(L.mrf_var_advance(_p(A), _p(B["med"]), n, _stream()), "mrf_var_advance")              # :650
        E = [float(e) for e in self.energies.get()]
        if pc is None:
            E[4] = 0.0
        energy = E[0] + E[1] + self.l1 * E[2] + self.l2 * E[3] + self.lc * E[4]                      # :564
        self.last = dict(op=op, da_raw=da, info=info, medians=self.med, E=tuple(E))
        return energy / (h * w)
