"""The reference-signature host wrappers (mrflow_b200.pipeline.*, .utils.*, .structure_refinement.*)
against the CPU oracle: same inputs, same call shapes as the reference's own callers (mrflow.py:98-140).

Tolerances: mu is a deterministic tree sum on the GPU and a pairwise sum in numpy (1e-16 relative
apart), so everything downstream of mu is compared to 1e-12 relative; flows to 1e-4 px (north_star);
discrete outputs exactly, away from decision boundaries."""
import argparse

import numpy as np
import pytest

from conftest import golden_triplet

pytestmark = pytest.mark.gpu


def _params(**kw):
    d = dict(debug_use_rigidity_gt=0, debug_save_frames=0, debug_compute_figure=0, rigidity_sigma_direction=1.0,
             rigidity_weight_cnn=0.25, rigidity_sigma_structure=1.0, rigidity_use_structure=1, scale_structure=1,
             nonlinear_structure_initialization=0, occlusion_reasoning=1, tempdir=".")
    d.update(kw)
    return argparse.Namespace(**d)


def _stub_mrf(I, unaries, lambd):
    """stand-in for the C++ solver (the unchanged consumer): per-pixel decision, no smoothing"""
    return (unaries[:, :, 1] > unaries[:, :, 0]).astype(np.int32)


@pytest.fixture(scope="module", params=["outside", "inside", "example", "moving"])
def trip(request):
    return golden_triplet(request.param)


def test_flow_homography(trip):
    from mrflow_b200.utils import flow_homography as fh
    from oracle import hotpath as hp
    t = trip
    h, w = t["shape"]
    y, x = np.mgrid[:h, :w]
    H = t["homographies"][1]
    ur, vr = fh.get_residual_flow(x, y, t["flow"][1][0], t["flow"][1][1], H)
    er, es = hp.get_residual_flow(x, y, t["flow"][1][0], t["flow"][1][1], H)
    assert np.array_equal(ur, er) and np.array_equal(vr, es) and ur.dtype == np.float32
    u, v = fh.get_full_flow(ur, vr, H, shape=(h, w))
    eu, ev = hp.get_full_flow(er, es, H, shape=(h, w))
    assert np.array_equal(u, eu) and np.array_equal(v, ev)
    u2, v2 = fh.get_full_flow(ur.astype(np.float64) * 1.37, vr.astype(np.float64), H, x=x, y=y)
    eu2, ev2 = hp.get_full_flow(er.astype(np.float64) * 1.37, es.astype(np.float64), H, x=x, y=y)
    assert np.array_equal(u2, eu2) and np.array_equal(v2, ev2)
    # round trip identity of SURVEY.md section 4
    assert np.abs(u - t["flow"][1][0]).max() < 1e-3
    import cv2
    hu, hv = fh.homography2flow(x, y, H)
    p = np.c_[x.ravel(), y.ravel()].astype("float32")
    uv = cv2.perspectiveTransform(p[:, None, :], H).squeeze() - p
    assert np.array_equal(hu, uv[:, 0].reshape(h, w)) and np.array_equal(hv, uv[:, 1].reshape(h, w))
    with pytest.raises(ValueError):
        fh.get_residual_flow(x + 1, y, t["flow"][1][0], t["flow"][1][1], H)


def test_robust_functions(golden):
    from mrflow_b200.utils import robust_functions as rf
    _, g = golden
    z = g["rho_x"]
    gens = dict(lor_fixed=rf.generate_lorentzian(1.5), lor_auto=rf.generate_lorentzian(),
                charb_fixed=rf.generate_charbonnier(1e-3), charb_auto=rf.generate_charbonnier(),
                charb_a=rf.generate_charbonnier(0.01, 0.45), gm_fixed=rf.generate_geman_mcclure(2.0),
                gm_auto=rf.generate_geman_mcclure(), quad=rf.generate_quadratic())
    for k, (f, df) in gens.items():
        np.testing.assert_allclose(f(z), g["rho_" + k], rtol=1e-13, atol=0, err_msg=k)
        np.testing.assert_allclose(df(z), g["drho_" + k], rtol=1e-13, atol=0, err_msg=k)
    # IEEE-only routes are bit-exact against the reference's outputs
    for k in ("charb_fixed", "gm_fixed", "quad"):
        assert np.array_equal(gens[k][0](z), g["rho_" + k])
    # 2-D input keeps its shape (optimize_structure_single_scale.py:289 passes images)
    assert gens["lor_fixed"][0](z[:4000].reshape(40, 100)).shape == (40, 100)


def test_interpolate_through_H(golden):
    from mrflow_b200.structure_refinement.interpolate_through_H import interpolate_through_H
    tag, g = golden
    t = golden_triplet(tag)
    J, m = interpolate_through_H(g["ith_I"], t["homographies"][1], g["ith_xn"], g["ith_yn"], compute_derivs=False)
    assert np.array_equal(J, g["ith_J"]) and np.array_equal(m, g["ith_mask"])
    assert J.dtype == np.float32 and m.dtype == np.float64
    dI = np.gradient(g["ith_I"])
    J2, m2, dx, dy = interpolate_through_H(g["ith_I"], t["homographies"][1], g["ith_xn"], g["ith_yn"],
                                           compute_derivs=True, dIdx=dI[1], dIdy=dI[0])
    from oracle import hotpath as hp
    ex, _ = hp.interpolate_through_H(dI[1], t["homographies"][1], g["ith_xn"], g["ith_yn"])
    assert np.array_equal(J2, g["ith_J"]) and np.array_equal(dx, ex)
    # single-channel image, identity: integer grid returns the image
    I1 = g["ith_I"][..., 0]
    y, x = np.mgrid[:I1.shape[0], :I1.shape[1]]
    J1, m1 = interpolate_through_H(I1, np.eye(3), x.astype(float), y.astype(float), compute_derivs=False)
    assert np.array_equal(J1, I1.astype(np.float32)) and m1.min() == 1.0


def test_structure2flow(golden):
    from mrflow_b200.pipeline import structure2flow as s2f
    tag, g = golden
    t = golden_triplet(tag)
    mu = list(g["mu"])
    u, v = s2f.structure_to_flow(g["structure1"], t["epipoles"][1], mu[1], t["homographies"][1], 0.41, _params())
    assert np.array_equal(u, g["s2f_u"], equal_nan=True) and np.array_equal(v, g["s2f_v"], equal_nan=True)
    rig = (t["rigidity"] > 0.5).astype(np.float64)
    r, cu, cv = s2f.combine_flow(g["structure1"], t["flow"][1], rig, t["epipoles"], mu, t["homographies"],
                                 [-0.37, 0.41], t["images"], _params())
    assert r is rig or np.array_equal(r, rig)
    assert np.array_equal(cu, g["combine_u"], equal_nan=True) and np.array_equal(cv, g["combine_v"], equal_nan=True)


@pytest.mark.parametrize("scale", [1, 0])
def test_compute_structure(trip, scale):
    from mrflow_b200.pipeline.compute_structure import compute_structure
    from oracle import hotpath as hp
    t = trip
    h, w = t["shape"]
    rs = np.random.RandomState(21)
    occ = [rs.rand(h, w) < 0.04, rs.rand(h, w) < 0.04]
    want = hp.compute_structure(t["images"], t["flow"], t["rigidity"], occ, [H.copy() for H in t["homographies"]],
                                t["epipoles"], hp.Params(scale_structure=scale), infer_mrf=_stub_mrf)
    got = compute_structure(t["images"], t["flow"], t["rigidity"], occ, [H.copy() for H in t["homographies"]],
                            t["epipoles"], _params(scale_structure=scale), infer_mrf=_stub_mrf)
    assert len(got) == 7
    S, Hs, mu, B, q, rig, oc = got
    np.testing.assert_allclose(mu, want[2], rtol=1e-14)
    np.testing.assert_allclose(B, want[3], rtol=1e-11)
    for f in range(2):
        np.testing.assert_allclose(q[f], want[4][f], rtol=0, atol=1e-9)
        assert S[f].dtype == np.float64 and S[f].shape == (h, w)
        fin = np.isfinite(want[0][f])
        assert np.array_equal(np.isfinite(S[f]), fin)
        np.testing.assert_allclose(S[f][fin], want[0][f][fin], rtol=1e-10)
        assert np.array_equal(Hs[f], t["homographies"][f])
    assert rig.dtype == bool and (rig == want[5]).mean() > 0.999
    assert oc is occ


def test_compute_structure_gates(trip):
    from mrflow_b200.pipeline.compute_structure import compute_structure
    t = trip
    h, w = t["shape"]
    occ = [np.zeros((h, w), bool)] * 2
    # all pixels non-rigid by CNN and direction -> TooMuchNonRigid (compute_structure.py:164-165)
    bad_flow = [[(np.random.RandomState(1).standard_normal((h, w)) * 20).astype(np.float32) for _ in range(2)] for _ in range(2)]
    with pytest.raises(Exception, match="TooMuchNonRigid"):
        compute_structure(t["images"], bad_flow, np.zeros((h, w)), occ, t["homographies"], t["epipoles"], _params(),
                          infer_mrf=_stub_mrf)
    with pytest.raises(Exception, match="TooFewStructureMatches"):
        compute_structure(t["images"], t["flow"], np.full((h, w), 0.9), occ, t["homographies"], t["epipoles"], _params(),
                          infer_mrf=lambda I, un, lam: np.zeros((h, w)))


@pytest.mark.parametrize("mask_mode", ["as_written", "rigid_only"])
def test_build_cost_volume(trip, mask_mode):
    from mrflow_b200.pipeline.build_cost_volume import build_cost_volume
    from oracle import hotpath as hp
    from test_gpu_kernels import _check_argmin, _check_costs
    t = trip
    h, w = t["shape"]
    rig = t["rigidity"].copy()
    rig[5:20, 60:80] = 0.0                      # some non-rigid pixels so both range rules have data
    want = hp.build_cost_volume(t["images"], t["flow"], rig, t["homographies"], t["epipoles"], hp.Params(),
                                range_mask=mask_mode)
    got = build_cost_volume(t["images"], t["flow"], rig, t["homographies"], t["epipoles"], _params(),
                            range_mask=mask_mode)
    assert len(got) == 6
    np.testing.assert_allclose(got[2], want[2], rtol=1e-14)          # mu
    np.testing.assert_allclose(got[3], want[3], rtol=1e-11)          # B
    assert got[3][1] == 1.0
    np.testing.assert_allclose(got[5], want[5], rtol=1e-10)          # labels
    assert got[5].shape == (51,)
    # The label set and mu carry the 1e-16 reduction-order difference into the sample coordinates, so
    # a handful of samples sit in a different 1/32-px bin than the oracle's; evaluate the oracle at
    # OUR labels / mu / B (bit-exact inputs) for the cost comparison itself.
    ch = [hp.prepare_channels(I) for I in t["images"]]
    noocc = [np.zeros((h, w), bool)] * 2
    cost, _, _ = hp.cost_volume(ch, rig, noocc, t["homographies"], got[3], got[2], got[4], got[5])
    for f in range(2):
        assert got[0][f].shape == (51, h, w) and got[0][f].dtype == np.float32
        _check_costs(got[0][f], cost[f], "frame %d" % f)
        fin = np.isfinite(want[1][f])
        np.testing.assert_allclose(got[1][f][fin], want[1][f][fin], rtol=1e-10)
    # and against the oracle's own volume: all but a vanishing fraction of samples agree
    close = np.isclose(np.stack(got[0]), np.stack(want[0]), rtol=1e-5, atol=1e-7)
    print("build_cost_volume[%s]: %.6f of the samples within 1e-5 of the oracle's own volume (own mu / B / labels)"
          % (mask_mode, close.mean()))
    assert close.mean() > 0.9999

    dev = build_cost_volume(t["images"], t["flow"], rig, t["homographies"], t["epipoles"], _params(),
                            range_mask=mask_mode, return_device=True, layout="HWL_I32")
    ci = dev[0][1].cpu().numpy()
    assert ci.shape == (h, w, 51) and ci.dtype == np.int32
    assert np.array_equal(ci, (got[0][1] * np.float32(1000.0)).astype(np.int32).transpose(1, 2, 0))
    _check_argmin(dev[6][1][1].cpu().numpy(), cost[1])


def test_occlusions_from_consistency(golden):
    """rigidity/occlusions.py against the reference's own outputs (tests/golden)."""
    from mrflow_b200.rigidity.occlusions import computeOcclusionsFromConsistency
    tag, g = golden
    t = golden_triplet(tag)
    backflow = [[g["occ_backflow"][f][0], g["occ_backflow"][f][1]] for f in range(2)]
    ob, of, both = computeOcclusionsFromConsistency(t["flow"], backflow, threshold=2.0)
    assert ob.dtype == bool
    assert np.array_equal(ob, g["occ_b"]) and np.array_equal(of, g["occ_f"]) and np.array_equal(both, g["occ_both"])
    # flows that leave the frame: the invalid-region bookkeeping (:66-70)
    far = [[(t["flow"][f][0] + 40).astype(np.float32), t["flow"][f][1]] for f in range(2)]
    from oracle import hotpath as hp
    want = hp.compute_occlusions_from_consistency(far, backflow, threshold=2.0)
    got = computeOcclusionsFromConsistency(far, backflow, threshold=2.0)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_image_weights_and_unaries(golden):
    """rigidity/inference.py:113,140-167 against the reference's own outputs."""
    from mrflow_b200.rigidity import inference
    tag, g = golden
    t = golden_triplet(tag)
    wts = inference.get_image_weights(t["images"][1])
    for k, name in enumerate(("w_h", "w_v", "w_ne", "w_se")):
        assert wts[k].dtype == np.float32 and wts[k].shape == g[name].shape
        # the mean behind beta is a tree sum here and a pairwise sum in numpy (1e-16 apart): one fp32 ulp
        np.testing.assert_allclose(wts[k], g[name], rtol=2e-7, atol=0)
        assert (wts[k] == g[name]).mean() > 0.999
    gray = t["images"][1][..., 0].copy()
    w1 = inference.get_image_weights(gray)
    from oracle import hotpath as hp
    for a, b in zip(w1, hp.get_image_weights(gray)):
        np.testing.assert_allclose(a, b, rtol=2e-7, atol=0)
    un = np.random.RandomState(3).rand(5, 7, 2)
    assert np.array_equal(inference.pack_unaries(un), hp.pack_unaries(un))


def test_compute_derivs_through_H(golden):
    """structure_refinement/interpolate_through_H.py:53-97 against the reference's own output."""
    from mrflow_b200.structure_refinement.interpolate_through_H import compute_derivs_through_H, interpolate_through_H
    from oracle import hotpath as hp
    tag, g = golden
    t = golden_triplet(tag)
    ch2 = hp.prepare_channels(t["images"][2])
    dx, dy = compute_derivs_through_H(ch2, t["homographies"][1])
    assert dx.dtype == np.float64 and dx.shape == ch2.shape
    assert np.array_equal(dx, g["derivs_dx"]) and np.array_equal(dy, g["derivs_dy"])
    # identity homography, single channel: central differences with replicated borders
    I = ch2[..., 0].copy()
    ex, ey = hp.compute_derivs_through_H(I, np.eye(3))
    gx, gy = compute_derivs_through_H(I, np.eye(3))
    assert np.array_equal(gx, ex) and np.array_equal(gy, ey)
    # compute_derivs=True without supplied derivative images now goes through the kernel as well
    h, w = I.shape
    y, x = np.mgrid[:h, :w]
    rs = np.random.RandomState(4)
    xn, yn = x + rs.uniform(-3, 3, (h, w)), y + rs.uniform(-3, 3, (h, w))
    got = interpolate_through_H(ch2, t["homographies"][1], xn, yn, compute_derivs=True)
    want = hp.interpolate_through_H_derivs(ch2, t["homographies"][1], xn, yn)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_compute_data_term(golden):
    """computeDataTerm (optimize_structure_single_scale.py:186-305) against the reference's own linear
    system: diag(A_data), b_data and E_data on the golden triplets, auto-sigma Lorentzian (:333)."""
    from mrflow_b200.structure_refinement.optimize_structure_single_scale import computeDataTerm
    from mrflow_b200.utils import robust_functions as rf
    from oracle import hotpath as hp
    tag, g = golden
    t = golden_triplet(tag)
    h, w = t["shape"]
    ch = [hp.prepare_channels(I) for I in t["images"]]
    cw = np.array([0.01, 1.0, 1.0])
    ev, dv = rf.generate_lorentzian()
    for f, (B, nb) in enumerate(((-0.37, 0), (0.41, 2))):
        A_data, b_data, E = computeDataTerm(g["dataterm_A%d" % f], ch[1], ch[nb], g["dataterm_rig"], g["dataterm_occ"],
                                            t["homographies"][f], B, g["mu"][f], t["epipoles"][f], ev, dv, 0.0, cw)
        assert A_data.shape == (h * w, h * w) and b_data.shape == (h * w,)
        diag = np.asarray(A_data.diagonal()).reshape(h, w)
        # sigma comes from an exact median (bit-identical); psi then involves only IEEE operations, so the
        # system matches to the last bit except where the 1e-16 effects of ... nothing: assert tightly
        np.testing.assert_allclose(diag, g["lin_diag%d" % f], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(b_data.reshape(h, w), g["lin_b%d" % f], rtol=1e-12, atol=1e-300)
        assert (diag == g["lin_diag%d" % f]).mean() > 0.99
        np.testing.assert_allclose(E, float(g["lin_E%d" % f]), rtol=1e-12)
    # fixed-sigma variant against the oracle, and a foreign closure is rejected
    ev2, dv2 = rf.generate_lorentzian(1.0)
    A2, b2, E2 = computeDataTerm(g["dataterm_A1"], ch[1], ch[2], g["dataterm_rig"], g["dataterm_occ"], t["homographies"][1],
                                 0.41, g["mu"][1], t["epipoles"][1], ev2, dv2, 0.0, cw)
    d_ref, b_ref, E_ref = hp.compute_data_term(g["dataterm_A1"], ch[1], ch[2], g["dataterm_rig"], g["dataterm_occ"],
                                               t["homographies"][1], 0.41, g["mu"][1], t["epipoles"][1], sigma=1.0)
    np.testing.assert_allclose(np.asarray(A2.diagonal()).reshape(h, w), d_ref, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(b2.reshape(h, w), b_ref, rtol=1e-12, atol=1e-300)
    with pytest.raises(TypeError):
        computeDataTerm(g["dataterm_A1"], ch[1], ch[2], g["dataterm_rig"], g["dataterm_occ"], t["homographies"][1], 0.41,
                        g["mu"][1], t["epipoles"][1], lambda x: x, lambda x: x, 0.0, cw)


def test_median5x5(golden):
    import torch
    from scipy import ndimage
    from mrflow_b200 import _lib
    import ctypes as C
    rs = np.random.RandomState(8)
    for shape in ((37, 53), (5, 9), (3, 4)):
        a = rs.standard_normal(shape)
        a[rs.rand(*shape) < 0.3] = 0.0                      # ties
        src = torch.from_numpy(a).cuda()
        dst = torch.empty_like(src)
        _lib.check(_lib.lib().mrf_median5x5_f64(C.c_void_p(src.data_ptr()), shape[0], shape[1], C.c_void_p(dst.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "median")
        assert np.array_equal(dst.cpu().numpy(), ndimage.median_filter(a, size=5))


def test_compute_structure_against_reference_end_to_end(golden, monkeypatch):
    """mrflow_b200.pipeline.compute_structure against the REFERENCE's compute_structure run end to end
    (tests/golden, generated with the reference's own C++ MRF solver), the same solver serving as the
    unchanged consumer here (oracle/_ref, prebuilt)."""
    from mrflow_b200.pipeline.compute_structure import compute_structure
    from oracle import trws
    tag, g = golden
    if "full_S" not in g:
        pytest.skip("no end-to-end fixture for this case")
    if not trws.available():
        pytest.skip("oracle/_ref not built")
    t = golden_triplet(tag)
    if tag == "inside":
        from mrflow_b200 import planeparallax as pp
        from test_oracle_golden import pin_epipoles
        pin_epipoles(monkeypatch, pp, "epipole_from_window", t, g)
    occ = [g["full_occ"][0], g["full_occ"][1]]
    S, Hs, mu, B, q, rig, oc = compute_structure(t["images"], t["flow"], t["rigidity"], occ,
                                                 [H.copy() for H in t["homographies"]], t["epipoles"], _params(),
                                                 infer_mrf=trws.infer_mrf)
    np.testing.assert_allclose(mu, g["full_mu"], rtol=1e-14)
    np.testing.assert_allclose(B, g["full_B"], rtol=1e-11)
    assert np.array_equal(np.array(q), g["full_q"])
    for f in range(2):
        fin = np.isfinite(g["full_S"][f])
        np.testing.assert_allclose(S[f][fin], g["full_S"][f][fin], rtol=1e-10)
    assert np.array_equal(rig, g["full_rigidity"])          # the MRF labelling is identical


def test_triplet_stream_equals_single_calls(trip):
    """mrflow_b200.dataset.TripletStream (several triplets in flight, pinned result buffers) returns
    exactly what build_cost_volume + combine_flow return one call at a time."""
    from mrflow_b200 import dataset
    from mrflow_b200.pipeline import structure2flow as s2f
    from mrflow_b200.pipeline.build_cost_volume import build_cost_volume
    t = trip
    rig = t["rigidity"].copy(); rig[5:20, 60:80] = 0.0
    variants = []
    for k in range(4):
        fl = [[t["flow"][f][0] + np.float32(0.01 * k), t["flow"][f][1]] for f in range(2)]
        variants.append(dict(images=t["images"], flow=fl, rigidity=rig, homographies=t["homographies"], epipoles=t["epipoles"]))
    import copy
    got = [copy.deepcopy(r) for r in dataset.run_triplets(variants, depth=2, range_mask="rigid_only")]   # slot buffers are reused
    assert len(got) == 4
    for v, r in zip(variants, got):
        want = build_cost_volume(v["images"], v["flow"], v["rigidity"], v["homographies"], v["epipoles"], None,
                                 range_mask="rigid_only")
        for f in range(2):
            assert np.array_equal(r["cost"][f], want[0][f]) and np.array_equal(r["structure"][f], want[1][f], equal_nan=True)
        assert r["mu"] == list(want[2]) and r["B"] == list(want[3]) and np.array_equal(r["labels"], want[5])
        _, u, vv = s2f.combine_flow(want[1][1], v["flow"][1], v["rigidity"], want[4], want[2], v["homographies"], want[3])
        assert np.array_equal(r["u"], u, equal_nan=True) and np.array_equal(r["v"], vv, equal_nan=True)


def test_flow2parallax_and_structure_helpers(golden):
    """The small public helpers of pipeline/compute_structure.py (:25-33, :517-537) against the reference's outputs."""
    from mrflow_b200.pipeline import compute_structure as cs
    tag, g = golden
    t = golden_triplet(tag)
    for f, B in ((0, -0.37), (1, 0.41)):
        p, vec, d = cs.flow2parallax(g["u_res%d" % f], g["v_res%d" % f], t["epipoles"][f])
        assert np.array_equal(p, g["parallax%d" % f]) and np.array_equal(d, g["dists%d" % f]) and vec.shape == p.shape + (2,)
        S = cs.parallax2normedstructure(p, t["epipoles"][f], B, g["mu"][f])
        assert np.array_equal(S, g["structure%d" % f], equal_nan=True)


def test_reference_chain_stages_4_to_6(golden, monkeypatch):
    """BASELINE.json configs[0] chain -- compute_structure -> optimize(override_optimization=1) -> combine_flow
    (mrflow.py:98-140) -- through the drop-in modules, against the REFERENCE's own outputs of the same chain
    (tests/golden; the reference's C++ MRF solver on both sides)."""
    from mrflow_b200.pipeline import compute_structure as cs, optimize_variational_refinement as ovr, structure2flow as s2f
    from oracle import trws
    tag, g = golden
    if "full_S" not in g:
        pytest.skip("no end-to-end fixture for this case")
    if not trws.available():
        pytest.skip("oracle/_ref not built")
    t = golden_triplet(tag)
    if tag == "inside":
        from mrflow_b200 import planeparallax as pp
        from test_oracle_golden import pin_epipoles
        pin_epipoles(monkeypatch, pp, "epipole_from_window", t, g)
    occ = [g["full_occ"][0], g["full_occ"][1]]
    params = _params(override_optimization=1)
    S, Hs, mu, B, q, rig, oc = cs.compute_structure(t["images"], t["flow"], t["rigidity"], occ,
                                                    [H.copy() for H in t["homographies"]], t["epipoles"], params,
                                                    infer_mrf=trws.infer_mrf)
    for reasoning in (1, 0):
        params.occlusion_reasoning = reasoning
        merged = ovr.optimize(t["images"], S, rig, q, Hs, B, mu, oc, params)[0]
        want = g["full_merged%d" % reasoning]
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(merged), fin)
        np.testing.assert_allclose(merged[fin], want[fin], rtol=1e-10)
    params.occlusion_reasoning = 1
    merged = ovr.optimize(t["images"], S, rig, q, Hs, B, mu, oc, params)[0]
    _, u, v = s2f.combine_flow(merged, t["flow"][1], rig, q, mu, Hs, B, t["images"], params)
    fin = np.isfinite(g["full_u"])
    assert np.abs(u - g["full_u"])[fin].max() <= 1e-4 and np.abs(v - g["full_v"])[fin].max() <= 1e-4    # north_star: 1e-4 px
    # with the reference's own stage-4 outputs as input, stages 5 and 6 are bit-exact
    params.occlusion_reasoning = 1
    m2 = ovr.optimize(t["images"], [g["full_S"][0], g["full_S"][1]], g["full_rigidity"], g["full_q"], Hs, g["full_B"],
                      g["full_mu"], occ, params)[0]
    assert np.array_equal(m2, g["full_merged1"], equal_nan=True)
    _, u2, v2 = s2f.combine_flow(m2, t["flow"][1], g["full_rigidity"], g["full_q"], g["full_mu"], Hs, g["full_B"],
                                 t["images"], params)
    assert np.array_equal(u2, g["full_u"], equal_nan=True) and np.array_equal(v2, g["full_v"], equal_nan=True)


def test_full_chain_sintel_size():
    """The chain of mrflow.py:98-140 at BASELINE.json's Sintel shape: drop-in modules on the GPU against the
    oracle on the CPU, the reference's C++ MRF solver as the consumer on both sides."""
    from mrflow_b200 import synth
    from mrflow_b200.pipeline import compute_structure as cs, optimize_variational_refinement as ovr, structure2flow as s2f
    from oracle import hotpath as hp, trws
    if not trws.available():
        pytest.skip("oracle/_ref not built")
    t = synth.make_triplet("sintel", seed=1003, flow_noise=0.03)
    h, w = t["shape"]
    rs = np.random.RandomState(9)
    occ = [rs.rand(h, w) < 0.03, rs.rand(h, w) < 0.03]
    want = hp.compute_structure(t["images"], t["flow"], t["rigidity"], occ, [H.copy() for H in t["homographies"]],
                                t["epipoles"], hp.Params(), infer_mrf=trws.infer_mrf)
    wm = hp.optimize_override(want[0], want[2], occ)
    _, wu, wv = hp.combine_flow(wm, t["flow"][1], want[5], want[4], want[2], want[1], want[3])
    params = _params(override_optimization=1)
    S, Hs, mu, B, q, rig, oc = cs.compute_structure(t["images"], t["flow"], t["rigidity"], occ,
                                                    [H.copy() for H in t["homographies"]], t["epipoles"], params,
                                                    infer_mrf=trws.infer_mrf)
    np.testing.assert_allclose(mu, want[2], rtol=1e-13)
    np.testing.assert_allclose(B, want[3], rtol=1e-10)
    assert (rig == want[5]).mean() > 0.9999                     # int32 unaries can flip at an integer boundary
    merged = ovr.optimize(t["images"], S, rig, q, Hs, B, mu, oc, params)[0]
    _, u, v = s2f.combine_flow(merged, t["flow"][1], rig, q, mu, Hs, B, t["images"], params)
    same = rig == want[5]
    fin = np.isfinite(wu) & np.isfinite(u) & same
    assert fin.mean() > 0.99
    assert np.abs(u - wu)[fin].max() <= 1e-4 and np.abs(v - wv)[fin].max() <= 1e-4          # north_star: 1e-4 px
    # and the chain reproduces the input flow on rigid pixels (SURVEY.md section 4 identity)
    rigid = rig & (~occ[0]) & (~occ[1])
    assert np.median(np.abs(u - t["flow"][1][0])[rigid]) < 0.05


def test_config1_full_size_against_reference():
    """BASELINE.json configs[0] at full size through the drop-in modules: the reference's example frames
    (1024x436), synthetic flows, uniform rigidity, override_optimization=1 -- compute_structure (reference MRF
    solver as the unchanged consumer) -> optimize -> combine_flow against the REFERENCE's own outputs
    (tests/golden/ref_example_full.npz: crops, and the hash of the rigidity map)."""
    from mrflow_b200.pipeline import compute_structure as cs, optimize_variational_refinement as ovr, structure2flow as s2f
    from oracle import trws
    from test_oracle_golden import _full_example
    if not trws.available():
        pytest.skip("oracle/_ref not built")
    mg, g, t = _full_example()
    params = _params(override_optimization=1, rigidity_weight_cnn=float(g["weight_cnn"]))
    S, Hs, mu, B, q, rig, oc = cs.compute_structure(t["images"], t["flow"], t["rigidity"], t["occlusions"],
                                                    [H.copy() for H in t["homographies"]], t["epipoles"], params,
                                                    infer_mrf=trws.infer_mrf)
    merged = ovr.optimize(t["images"], S, rig, q, Hs, B, mu, oc, params)[0]
    _, u, v = s2f.combine_flow(merged, t["flow"][1], rig, q, mu, Hs, B, t["images"], params)
    np.testing.assert_allclose(mu, g["crop_mu"], rtol=1e-14)
    np.testing.assert_allclose(B, g["crop_B"], rtol=1e-11)
    assert np.array_equal(np.array(q), g["crop_q"])
    assert mg.sha(np.ascontiguousarray(np.asarray(rig))) == str(g["sha_rigidity"])      # the MRF labelling is identical
    c = mg.CROP
    for name, got in (("S0", S[0]), ("S1", S[1]), ("merged", merged)):
        want = g["crop_" + name]
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[c][fin], want[fin], rtol=1e-10)
    assert np.abs(u[c] - g["crop_u"]).max() <= 1e-4 and np.abs(v[c] - g["crop_v"]).max() <= 1e-4   # north_star: 1e-4 px
    assert u.dtype == np.float32 and u.shape == (436, 1024)


def test_build_cost_volume_auto_sigma(trip):
    """build_cost_volume with the reference's auto-sigma Lorentzian (sigma=None; utils/robust_functions.py:55-58)
    against the oracle's same mode, at the GPU's own scalars (see test_build_cost_volume)."""
    from mrflow_b200.pipeline.build_cost_volume import build_cost_volume
    from oracle import hotpath as hp
    from test_gpu_kernels import _check_costs
    t = trip
    h, w = t["shape"]
    got = build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], _params(), rho="lorentzian",
                            sigma=None, range_mask="rigid_only")
    ch = [hp.prepare_channels(I) for I in t["images"]]
    noocc = [np.zeros((h, w), bool)] * 2
    want, _, _ = hp.cost_volume(ch, t["rigidity"], noocc, t["homographies"], got[3], got[2], got[4], got[5], kind="lorentzian",
                                sigma=0.0, labels=[0, 13, 25, 38, 50])
    for f in range(2):
        assert got[0][f].shape == (51, h, w)
        _check_costs(got[0][f][[0, 13, 25, 38, 50]], want[f], "auto-sigma frame %d" % f)


@pytest.mark.parametrize("layout", ["LHW_F32", "HWL_I32"])
def test_pipelined_host_call_equals_resident_pass(trip, layout):
    """The plain build_cost_volume call (uploads, row-chunked kernels and pitched downloads overlapped over
    PCIe) returns exactly what the device-resident pass leaves in HBM, in both volume layouts."""
    from mrflow_b200.pipeline.build_cost_volume import build_cost_volume
    t = trip
    host = build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], _params(), layout=layout,
                             range_mask="rigid_only")
    dev = build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], _params(), layout=layout,
                            range_mask="rigid_only", return_device=True)
    for f in range(2):
        assert np.array_equal(host[0][f], dev[0][f].cpu().numpy())
        assert np.array_equal(host[1][f], dev[1][f].cpu().numpy(), equal_nan=True)
    assert host[2] == dev[2] and host[3] == dev[3] and np.array_equal(host[5], dev[5])
    assert all(np.array_equal(a, b) for a, b in zip(host[4], dev[4]))


@pytest.mark.gpu
def test_flo_decode_against_reference(tmp_path):
    """utils/flow_io.py:33-48: mrflow_b200.utils.flow_io.flow_read (the file goes to HBM as it lies on disk, the planes
    are split by mrf_flo_decode) against what the reference's own flow_read returned for the stored file, bit for bit;
    a round trip through flow_write at the Sintel size; the reference's two assertions."""
    import os
    import torch
    from conftest import GOLDEN
    from mrflow_b200.utils import flow_io
    g = np.load(os.path.join(GOLDEN, "ref_flo.npz"))
    name = str(tmp_path / "a.flo")
    g["flo"].tofile(name)
    u, v = flow_io.flow_read(name)
    assert u.dtype == np.float32 and u.shape == g["u"].shape
    assert np.array_equal(u.view(np.uint32), g["u"].view(np.uint32)) and np.array_equal(v.view(np.uint32), g["v"].view(np.uint32))
    ud, vd = flow_io.flow_read(name, device_out=True)
    assert ud.is_cuda and vd.is_cuda
    assert np.array_equal(ud.cpu().numpy().view(np.uint32), g["u"].view(np.uint32))
    rs = np.random.RandomState(5)
    U, V = rs.standard_normal((436, 1024)).astype(np.float32), rs.standard_normal((436, 1024)).astype(np.float32)
    big = str(tmp_path / "b.flo")
    flow_io.flow_write(big, U, V)
    u2, v2 = flow_io.flow_read(big)
    assert np.array_equal(u2, U) and np.array_equal(v2, V)
    from oracle import hotpath as hp
    u3, v3 = hp.flow_read(big)                                  # the writer's file is one the reference layout reads
    assert np.array_equal(u3, U) and np.array_equal(v3, V)
    bad = g["flo"].copy(); bad[0] ^= 1
    bad.tofile(name)
    with pytest.raises(AssertionError):
        flow_io.flow_read(name)
    short = g["flo"][:-8]
    short.tofile(name)
    with pytest.raises(AssertionError):
        flow_io.flow_read(name)
