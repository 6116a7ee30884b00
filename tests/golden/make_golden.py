#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on seeded inputs.

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Import shims (none alters arithmetic): sys.path entries for the
reference's Python-2 implicit relative imports, MRFLOW_HOME, a stub ``chumpy`` module (only
``refine_A`` / the epipole BFGS use it; both are bypassed below), ``np.float`` alias, stub
matplotlib-free ``utils.plot_debug`` / ``utils.compute_figure`` modules.

The fixtures are small (a 96x72 triplet, crops and checksums); whole volumes are not stored.
"""
import os
import sys
import types
import hashlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


class _Ch:
    """The handful of chumpy.Ch operations correct_epipole uses (pipeline/compute_structure.py:494-504), on a
    torch fp64 tensor with autograd standing in for chumpy's own automatic differentiation.  Lives in this
    generator only: it lets the reference's REAL correct_epipole run (objective, BFGS call, acceptance rule
    all untouched) so that its result can be stored and the restated analytic gradient checked against it."""
    __array_priority__ = 1000.0          # numpy defers ndarray (op) _Ch to the reflected method, as it does for chumpy

    def __init__(self, t):
        self.t = t

    @staticmethod
    def _t(o):
        import torch
        if isinstance(o, _Ch):
            return o.t
        return torch.as_tensor(np.asarray(o, dtype=np.float64))

    def __getitem__(self, i): return _Ch(self.t[i])
    def __add__(self, o): return _Ch(self.t + self._t(o))
    __radd__ = __add__
    def __sub__(self, o): return _Ch(self.t - self._t(o))
    def __rsub__(self, o): return _Ch(self._t(o) - self.t)
    def __mul__(self, o): return _Ch(self.t * self._t(o))
    __rmul__ = __mul__
    def __truediv__(self, o): return _Ch(self.t / self._t(o))
    def __rtruediv__(self, o): return _Ch(self._t(o) / self.t)
    def __pow__(self, e): return _Ch(self.t ** e)
    def sum(self): return _Ch(self.t.sum())
    def __call__(self): return self.t.detach().numpy().copy()

    def dr_wrt(self, wrt):
        import torch
        (g,) = torch.autograd.grad(self.t, wrt.t, retain_graph=True)
        return g.detach().numpy().reshape(1, -1)


def _ch_namespace():
    import torch
    ns = types.SimpleNamespace()
    ns.array = lambda a: _Ch(torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True))
    ns.maximum = lambda a, b: _Ch(torch.maximum(_Ch._t(a), _Ch._t(b)))
    ns.abs = lambda a: _Ch(torch.abs(_Ch._t(a)))
    return ns


def import_reference():
    os.environ["MRFLOW_HOME"] = REF
    for p in (REF, REF + "/structure_refinement", REF + "/utils"):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "float"):
        np.float = float
    ch_mod = types.ModuleType("chumpy")
    ch_mod.ch = _ch_namespace()          # only correct_epipole's few operations exist; refine_A stays disabled
    sys.modules.setdefault("chumpy", ch_mod)
    for name in ("utils.plot_debug", "utils.compute_figure", "utils.print_exception", "utils.spyder_debug"):
        m = types.ModuleType(name)
        m.__dict__["__all__"] = []
        sys.modules.setdefault(name, m)
    import utils  # noqa: F401  (reference package)
    for name in ("plot_debug", "compute_figure", "print_exception", "spyder_debug"):
        setattr(sys.modules["utils"], name, sys.modules["utils." + name])
    from utils import flow_homography, robust_functions
    from pipeline import compute_structure as cs
    from pipeline import structure2flow as s2f
    from rigidity import inference
    import interpolate_through_H as ith
    import optimize_structure_single_scale as oss
    return dict(fh=flow_homography, rf=robust_functions, cs=cs, s2f=s2f, inf=inference, ith=ith, oss=oss)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    sys.path.insert(0, ROOT)
    from mrflow_b200 import synth
    ref = import_reference()
    fh, rf, cs, s2f, inf, ith, oss = (ref[k] for k in ("fh", "rf", "cs", "s2f", "inf", "ith", "oss"))
    import cv2
    import argparse

    for tag, kw in (("outside", dict(shape=(72, 96), seed=7, epipole="outside", flow_noise=0.02)),
                    ("inside", dict(shape=(72, 96), seed=8, epipole="inside", moving_objects=True)),
                    ("example", dict(shape=(72, 96), seed=9, epipole="outside", flow_noise=0.05)),
                    ("moving", dict(shape=(72, 96), seed=11, epipole="outside", moving_objects=True, flow_noise=0.02))):
        t = synth.make_triplet(n_waves=8, **kw)
        example_images = None
        if tag == "example":
            # BASELINE.json configs[0]: the reference's own example frames (example/frame1-3.png, loaded as
            # load_data.py:52 does: BGR -> RGB, /255), a 96x72 crop, with synthetic flows
            example_images = [cv2.imread(REF + "/example/frame%d.png" % i)[164:236, 400:496, ::-1] / 255.0 for i in (1, 2, 3)]
            t["images"] = example_images
        h, w = t["shape"]
        y, x = np.mgrid[:h, :w]
        out = {}
        # --- a1 / a2 / a3 ------------------------------------------------------------------
        par, dists, res = [], [], []
        for f in range(2):
            ur, vr = fh.get_residual_flow(x, y, t["flow"][f][0], t["flow"][f][1], t["homographies"][f])
            p, fv, d = cs.flow2parallax(ur, vr, t["epipoles"][f])
            out["u_res%d" % f], out["v_res%d" % f] = ur, vr
            out["parallax%d" % f], out["dists%d" % f] = p, d
            par.append(p); dists.append(d); res.append((ur, vr))
        mu = [d.mean() for d in dists]
        out["mu"] = np.array(mu)
        for f, B in ((0, -0.37), (1, 0.41)):
            out["structure%d" % f] = cs.parallax2normedstructure(par[f], t["epipoles"][f], B, mu[f])
        # --- a6 ----------------------------------------------------------------------------
        params = argparse.Namespace(debug_compute_figure=0)
        u, v = s2f.structure_to_flow(out["structure1"], t["epipoles"][1], mu[1], t["homographies"][1], 0.41, params)
        out["s2f_u"], out["s2f_v"] = u, v
        rig01 = (t["rigidity"] > 0.5).astype(np.float64)
        _, cu, cv = s2f.combine_flow(out["structure1"], t["flow"][1], rig01, t["epipoles"], mu,
                                     t["homographies"], [-0.37, 0.41], t["images"], params)
        out["combine_u"], out["combine_v"] = cu, cv
        # --- a7 ----------------------------------------------------------------------------
        for f in range(2):
            out["unary%d" % f] = inf.get_unaries_rigid(x, y, res[f][0], res[f][1],
                                                       t["epipoles"][f][0], t["epipoles"][f][1], sigma=1.0)
        # --- a8 ----------------------------------------------------------------------------
        rs = np.random.RandomState(99)
        I = rs.rand(h, w, 3) * 255
        xn = x + rs.uniform(-6, 6, (h, w))
        yn = y + rs.uniform(-6, 6, (h, w))
        J, m = ith.interpolate_through_H(I, t["homographies"][1], xn, yn, compute_derivs=False)
        out["ith_I"], out["ith_xn"], out["ith_yn"], out["ith_J"], out["ith_mask"] = I, xn, yn, J, m
        # --- a10 ---------------------------------------------------------------------------
        z = (rs.standard_normal(4001) * 3) ** 2
        out["rho_x"] = z
        gens = dict(lor_fixed=rf.generate_lorentzian(1.5), lor_auto=rf.generate_lorentzian(),
                    charb_fixed=rf.generate_charbonnier(1e-3), charb_auto=rf.generate_charbonnier(),
                    charb_a=rf.generate_charbonnier(0.01, 0.45), gm_fixed=rf.generate_geman_mcclure(2.0),
                    gm_auto=rf.generate_geman_mcclure(), quad=rf.generate_quadratic())
        for k, (fn, dfn) in gens.items():
            out["rho_" + k], out["drho_" + k] = fn(z), dfn(z)
        # --- a11: channels exactly as optimize_structure_single_scale.py:350-371 builds them ----
        chans = []
        for Iim in t["images"]:
            g = cv2.cvtColor(Iim.astype("float32"), cv2.COLOR_RGB2GRAY).astype("float64") * 255.0
            gy, gx = np.gradient(g)
            gy = cv2.GaussianBlur(gy, ksize=(3, 3), sigmaX=-1)
            gx = cv2.GaussianBlur(gx, ksize=(3, 3), sigmaX=-1)
            chans.append(np.dstack((g, gy, gx)))
        cw = np.concatenate((0.01 * np.array([1.0]), [1.0, 1.0]))
        rig = rig01
        occ = (rs.rand(h, w) < 0.05).astype(np.float64)
        robust_ev, robust_deriv = rf.generate_lorentzian(1.0)
        # computeDataTerm needs scipy.ndimage / sparse names from its module namespace
        for f, (B, nb) in enumerate(((-0.37, 0), (0.41, 2))):
            A = out["structure1"] * (mu[0] / mu[1] if f == 0 else 1.0)
            A = np.nan_to_num(A, posinf=0.0, neginf=0.0)
            _, _, E = oss.computeDataTerm(A, chans[1], chans[nb], rig, occ, t["homographies"][f], B, mu[f],
                                          t["epipoles"][f], robust_ev, robust_deriv, 0.0, cw)
            out["dataterm_A%d" % f] = A
            out["dataterm_E%d" % f] = np.array(E)
            # the full linearised data term as the variational refinement uses it: auto-sigma Lorentzian
            # (optimize_structure_single_scale.py:333)
            ev_auto, dv_auto = rf.generate_lorentzian()
            A_data, b_data, E_auto = oss.computeDataTerm(A, chans[1], chans[nb], rig, occ, t["homographies"][f], B, mu[f],
                                                         t["epipoles"][f], ev_auto, dv_auto, 0.0, cw)
            out["lin_diag%d" % f] = np.asarray(A_data.diagonal()).reshape(h, w)
            out["lin_b%d" % f] = b_data.reshape(h, w)
            out["lin_E%d" % f] = np.array(E_auto)
            if f == 1:
                ddx, ddy = ith.compute_derivs_through_H(chans[nb], t["homographies"][f])
                out["derivs_dx"], out["derivs_dy"] = ddx, ddy
        out["dataterm_occ"] = occ
        out["dataterm_rig"] = rig
        # --- a12: per-label planes from the reference geometry + sampler + rho ----------------
        labels = np.linspace(-3.0, 4.5, 7)
        planes = np.zeros((2, len(labels), h, w), np.float32)
        for f, (B, nb) in enumerate(((-0.37, 0), (0.41, 2))):
            q = t["epipoles"][f]
            H = t["homographies"][f]
            for li, lab in enumerate(labels):
                a = lab * (mu[0] / mu[1] if f == 0 else 1.0)
                # geometry lines :210-235 evaluated through the reference function itself is not
                # separable from its linearisation, so replay them here verbatim in spirit with
                # the reference's own sampler and rho.
                vx = q[0] - x; vy = q[1] - y
                nrm = np.maximum(0.5, np.sqrt(vx ** 2 + vy ** 2))
                n = np.sqrt(vx ** 2 + vy ** 2) / mu[f]
                r = a * B * n / (B * a / mu[f] - 1)
                J, m = ith.interpolate_through_H(chans[nb], H, x + r * (vx / nrm), y + r * (vy / nrm),
                                                 compute_derivs=False)
                Iz = ((J - chans[1]) * cw[np.newaxis, np.newaxis, :]).mean(axis=2)
                Iz[m == 0] = 0; Iz[occ == 1] = 0; Iz[rig == 0] = 0
                planes[f, li] = robust_ev(Iz ** 2)
        out["cv_labels"] = labels
        out["cv_cost_sha"] = np.array(sha(planes))
        out["cv_cost_crop"] = planes[:, :, 20:36, 30:62].copy()
        out["cv_argmin"] = planes.argmin(axis=1).astype(np.int32)
        out["cv_min"] = planes.min(axis=1)
        # --- section 8f rows: occlusion maps (rigidity/occlusions.py) and MRF edge weights (inference.py:140)
        from rigidity import occlusions as occmod
        rs2 = np.random.RandomState(123)
        def smooth(scale):
            f = cv2.GaussianBlur(rs2.standard_normal((h, w)).astype(np.float32), (0, 0), 4.0) * scale
            return f.astype(np.float32)
        backflow = [[(-t["flow"][f][0] + smooth(6.0)).astype(np.float32), (-t["flow"][f][1] + smooth(6.0)).astype(np.float32)]
                    for f in range(2)]
        ob, of, oboth = occmod.computeOcclusionsFromConsistency(t["flow"], backflow, threshold=2.0)
        out["occ_backflow"] = np.array(backflow)
        out["occ_b"], out["occ_f"], out["occ_both"] = ob, of, oboth
        wts = inf.get_image_weights(t["images"][1])
        for k, name in enumerate(("w_h", "w_v", "w_ne", "w_se")):
            out[name] = wts[k]
        # --- a5: the reference's own correct_epipole (pipeline/compute_structure.py:453-511), its chumpy
        # derivative supplied by the torch-autograd shim above
        import contextlib, io
        rig_thr = cv2.erode((t["rigidity"] > 0.5).astype("uint8"), np.ones((3, 3))) > 0       # :44
        qs = []
        for f in range(2):
            with contextlib.redirect_stdout(io.StringIO()):
                qs.append(np.asarray(cs.correct_epipole(res[f][0], res[f][1], np.asarray(t["epipoles"][f], dtype=np.float64), rig_thr),
                                     dtype=np.float64))
        out["epi_rigid"] = rig_thr
        out["epi_q_new"] = np.array(qs)
        # --- the complete stage, reference code end to end: pipeline/compute_structure.py:37-449 with the
        # reference's own C++ MRF solver (oracle/_ref, built by oracle/build_ref.py); with the epipole inside the
        # frame its correct_epipole runs on the shimmed chumpy.
        if True:
            sys.path.insert(0, ROOT)
            from oracle import trws
            if trws.available():
                inf.eff_trws = trws._ext()
                inf.HAVE_TRWS = True
                occ2 = [rs2.rand(h, w) < 0.04, rs2.rand(h, w) < 0.04]
                p = argparse.Namespace(debug_use_rigidity_gt=0, debug_save_frames=0, debug_compute_figure=0,
                                       rigidity_sigma_direction=1.0, rigidity_weight_cnn=0.25, rigidity_sigma_structure=1.0,
                                       rigidity_use_structure=1, scale_structure=1, nonlinear_structure_initialization=0,
                                       occlusion_reasoning=1, tempdir=".")
                with contextlib.redirect_stdout(io.StringIO()):
                    full = cs.compute_structure(t["images"], t["flow"], t["rigidity"], occ2,
                                                [H.copy() for H in t["homographies"]], t["epipoles"], p)
                out["full_occ"] = np.array(occ2)
                out["full_S"] = np.array(full[0]); out["full_mu"] = np.array(full[2]); out["full_B"] = np.array(full[3])
                out["full_q"] = np.array(full[4]); out["full_rigidity"] = np.asarray(full[5])
                # stage 5 with --override_optimization 1 (pipeline/optimize_variational_refinement.py:20-73)
                # and stage 6 (pipeline/structure2flow.py:41-82): the rest of the reference chain of configs[0]
                from pipeline import optimize_variational_refinement as ovr
                p.override_optimization = 1
                for reasoning in (1, 0):
                    p.occlusion_reasoning = reasoning
                    merged = ovr.optimize(t["images"], full[0], full[5], full[4], full[1], full[3], full[2], occ2, p)[0]
                    out["full_merged%d" % reasoning] = merged
                p.occlusion_reasoning = 1
                _, fu, fv = s2f.combine_flow(out["full_merged1"], t["flow"][1], full[5], full[4], full[2], full[1], full[3],
                                             t["images"], p)
                out["full_u"], out["full_v"] = fu, fv
                # --- section 8f row 4: the variational refinement itself, optimize_structure_single_scale.py:310-652
                # (what stage 5 runs without --override_optimization; one scale, parameters.py:43-45).  Two
                # py2-isms are shimmed: dict.has_key on vars(params) and the tol= keyword of scipy's minres.
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from variational_case import CASES, LAMBDA_1, LAMBDA_2, variational_inputs, with_nonrigid
                from scipy.sparse import linalg as spla
                from oracle import variational as ov
                import construct_matrices as cm

                class HK(dict):
                    def has_key(self, k):
                        return k in self
                oss.vars = lambda p_: HK(p_.__dict__)
                oss.solve = lambda A_, b_, method, tol=1e-3: spla.minres(A_, b_, rtol=tol)[0]
                assert abs(cm.construct_gradient_matrix((h, w)) - ov.gradient_matrix((h, w))).max() == 0
                assert abs(cm.construct_2ndorder_matrix((h, w)) - ov.second_order_matrix((h, w))).max() == 0
                pv = argparse.Namespace(variational_dataterm_grayscale=1, variational_min_weight=0.0, debug_save_frames=0,
                                        tempdir=".")
                vin = variational_inputs(t, out)
                for name, lc, n_outer, masked in CASES:
                    src = with_nonrigid(vin) if masked else vin
                    kwv = {k: (v.copy() if isinstance(v, np.ndarray) else [np.array(e).copy() for e in v]) for k, v in src.items()}
                    A_ref, en = oss.optimizeStructureSingleScale(
                        kwv["A_init"], kwv["A_bwd_reference"], kwv["A_fwd_reference"], kwv["occlusions_bwd"],
                        kwv["occlusions_fwd"], kwv["occlusions_both"], kwv["rigidity"], kwv["images"], kwv["H_ar"],
                        [float(b) for b in kwv["B_ar"]], kwv["q_ar"], [float(m) for m in kwv["mu_ar"]], LAMBDA_1, LAMBDA_2, lc,
                        N_outer=n_outer, params=pv)
                    out["var_A_" + name] = A_ref
                    out["var_E_" + name] = np.array(en)
                # --- the reference's default chain: stage 5 WITH the variational refinement (parameters.py:38-47:
                # one scale, 5 outer iterations) and stage 6 on its result
                import optimize_structure_multi_scale as osm
                osm.optimize_structure_single_scale = oss
                ovr.optimize_structure_multi_scale = osm
                p.override_optimization = 0
                p.occlusion_reasoning = 1
                for k_, v_ in dict(variational_lambda_1storder=0.5, variational_lambda_2ndorder=5.0, variational_lambda_consistency=0.0,
                                   variational_N_outer=5, variational_N_inner=1, variational_scale_factor=0.5, variational_N_scales=1,
                                   variational_last_scale=0, variational_dataterm_grayscale=1, variational_min_weight=0.0).items():
                    setattr(p, k_, v_)
                refined = ovr.optimize([I.copy() for I in t["images"]], full[0], full[5], full[4], full[1], full[3], full[2], occ2, p)[0]
                out["full_refined"] = refined
                _, fu, fv = s2f.combine_flow(refined, t["flow"][1], full[5], full[4], full[2], full[1], full[3], t["images"], p)
                out["full_refined_u"], out["full_refined_v"] = fu, fv
                # a two-level pyramid (optimize_structure_multi_scale.py:52-143: resampling of images, structures and
                # masks, scaled H / B / q / mu per level), two outer iterations per level
                p.variational_N_scales = 2
                p.variational_N_outer = 2
                out["full_refined_ms2"] = ovr.optimize([I.copy() for I in t["images"]], full[0], full[5], full[4], full[1], full[3],
                                                       full[2], occ2, p)[0]
        if example_images is not None:
            out["images"] = np.stack(example_images)
        np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % tag), **out)
        print(tag, "written:", len(out), "arrays")


FULL_KEYS = ("S0", "S1", "mu", "B", "q", "rigidity", "merged", "u", "v")
CROP = (slice(200, 232), slice(480, 544))          # 32 x 64


def full_example_inputs(images_u8, occ_bits):
    """The inputs of the full-size configs[0] chain, rebuilt identically by the generator and by the tests: the
    reference's example frames (uint8 RGB from the fixture, /255 as load_data.py:52), synthetic flows /
    homographies / epipoles of a seeded Sintel-shape triplet, the uniform 0.6 rigidity prior of --no_init
    (load_data.py:153-155) and the stored occlusion maps."""
    from mrflow_b200 import synth
    t = synth.make_triplet("sintel", seed=1005, flow_noise=0.03)
    h, w = t["shape"]
    t["images"] = [images_u8[i].astype(np.float64) / 255.0 for i in range(3)]
    t["rigidity"] = np.full((h, w), 0.6)
    t["occlusions"] = [np.unpackbits(occ_bits[i])[:h * w].reshape(h, w).astype(bool) for i in range(2)]
    return t


def main_full():
    """BASELINE.json configs[0] at FULL size: example/frame1-3.png (1024x436) with synthetic precomputed flows, uniform
    rigidity, --override_optimization 1 -- the reference's own chain of mrflow.py:98-140 (compute_structure with its
    C++ MRF solver -> optimize -> combine_flow).  Stored: the frames (uint8), the occlusion maps (bit-packed), and of
    every output its sha256 and a 32x64 crop."""
    sys.path.insert(0, ROOT)
    import argparse
    import contextlib
    import io
    import cv2
    from mrflow_b200 import synth
    ref = import_reference()
    cs, s2f, inf = ref["cs"], ref["s2f"], ref["inf"]
    from oracle import trws
    assert trws.available(), "build oracle/_ref first (python -m oracle.build_ref)"
    inf.eff_trws = trws._ext()
    inf.HAVE_TRWS = True
    from rigidity import occlusions as occmod
    from pipeline import optimize_variational_refinement as ovr
    images_u8 = np.stack([cv2.imread(REF + "/example/frame%d.png" % i)[:, :, ::-1] for i in (1, 2, 3)])   # BGR -> RGB, load_data.py:52
    t0 = synth.make_triplet("sintel", seed=1005, flow_noise=0.03)
    h, w = t0["shape"]
    rs = np.random.RandomState(321)

    def smooth(scale):
        return (cv2.GaussianBlur(rs.standard_normal((h, w)).astype(np.float32), (0, 0), 6.0) * scale).astype(np.float32)
    backflow = [[(-t0["flow"][f][0] + smooth(9.0)).astype(np.float32), (-t0["flow"][f][1] + smooth(9.0)).astype(np.float32)]
                for f in range(2)]
    ob, of, _ = occmod.computeOcclusionsFromConsistency(t0["flow"], backflow, threshold=3.0)
    occ_bits = np.stack([np.packbits(np.asarray(o).astype(np.uint8).ravel()) for o in (ob, of)])
    t = full_example_inputs(images_u8, occ_bits)
    p = argparse.Namespace(debug_use_rigidity_gt=0, debug_save_frames=0, debug_compute_figure=0, rigidity_sigma_direction=1.0,
                           rigidity_weight_cnn=0.0, rigidity_sigma_structure=1.0, rigidity_use_structure=1, scale_structure=1,
                           nonlinear_structure_initialization=0, occlusion_reasoning=1, tempdir=".", override_optimization=1)
    with contextlib.redirect_stdout(io.StringIO()):
        full = cs.compute_structure(t["images"], t["flow"], t["rigidity"], t["occlusions"], [H.copy() for H in t["homographies"]],
                                    t["epipoles"], p)
        merged = ovr.optimize(t["images"], full[0], full[5], full[4], full[1], full[3], full[2], t["occlusions"], p)[0]
        _, u, v = s2f.combine_flow(merged, t["flow"][1], full[5], full[4], full[2], full[1], full[3], t["images"], p)
    vals = dict(S0=full[0][0], S1=full[0][1], mu=np.array(full[2]), B=np.array(full[3]), q=np.array(full[4]),
                rigidity=np.asarray(full[5]), merged=merged, u=u, v=v)
    out = {"images_u8": images_u8, "occ_bits": occ_bits, "weight_cnn": np.array(0.0)}
    for k in FULL_KEYS:
        a = np.ascontiguousarray(vals[k])
        out["sha_" + k] = np.array(sha(a))
        out["crop_" + k] = a[CROP].copy() if a.shape == (h, w) else a.copy()
    out["occluded_fraction"] = np.array([float(np.mean(o)) for o in t["occlusions"]])
    out["rigid_fraction"] = np.array(float(np.mean(full[5])))
    np.savez_compressed(os.path.join(HERE, "ref_example_full.npz"), **out)
    print("example_full written:", len(out), "arrays; rigid fraction %.3f, occluded %s" % (out["rigid_fraction"], out["occluded_fraction"]))


if __name__ == "__main__":
    if "--full-only" not in sys.argv:
        main()
    main_full()
