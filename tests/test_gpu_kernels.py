"""Parity of the CUDA kernels (called through the C ABI) against the CPU oracle on the same seeded
inputs.  Bars (BASELINE.json north_star): geometry bit-exact (target 1e-4 px), fp32 costs within
1e-5 relative, argmin labels exact wherever the best-vs-second gap exceeds that tolerance."""
import numpy as np
import pytest

from conftest import golden_triplet

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-5     # north_star: fp32 costs within 1e-5 relative
COST_ATOL = 1e-7     # absolute floor for costs near zero (log(1+z) with z ~ 0)


@pytest.fixture(scope="module")
def dev():
    import torch
    from mrflow_b200 import device
    assert torch.cuda.is_available()
    return device


def cu(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.fixture(scope="module", params=["outside", "inside", "example", "moving"])
def trip(request):
    return golden_triplet(request.param)


def _oracle_front(t):
    from oracle import hotpath as hp
    h, w = t["shape"]
    y, x = np.mgrid[:h, :w]
    out = []
    for f in range(2):
        ur, vr = hp.get_residual_flow(x, y, t["flow"][f][0], t["flow"][f][1], t["homographies"][f])
        p, _, d = hp.flow2parallax(ur, vr, t["epipoles"][f])
        out.append((ur, vr, p, d))
    return out


def test_flow_to_parallax_and_mu(dev, trip):
    t = trip
    h, w = t["shape"]
    ref = _oracle_front(t)
    for f in range(2):
        Hinv = np.linalg.inv(t["homographies"][f])
        ur, vr, par, d = dev.flow_to_parallax(cu(t["flow"][f][0]), cu(t["flow"][f][1]), Hinv, t["epipoles"][f])
        assert eq(ur.cpu(), ref[f][0]) and eq(vr.cpu(), ref[f][1])
        assert eq(par.cpu(), ref[f][2]) and eq(d.cpu(), ref[f][3])
        s = float(dev.epipole_dist_sum(h, w, t["epipoles"][f]).cpu()[0])
        assert abs(s / (h * w) - ref[f][3].mean()) <= 1e-13 * ref[f][3].mean()


def test_parallax_structure_flow_roundtrip(dev, trip):
    from oracle import hotpath as hp
    t = trip
    ref = _oracle_front(t)
    for f, B in ((0, -0.37), (1, 0.41)):
        mu = ref[f][3].mean()
        S_ref = hp.parallax2normedstructure(ref[f][2], t["epipoles"][f], B, mu)
        S = dev.parallax_to_structure(cu(ref[f][2]), t["epipoles"][f], B, mu)
        assert eq(S.cpu(), S_ref)
        u_ref, v_ref = hp.structure_to_flow(S_ref, t["epipoles"][f], mu, t["homographies"][f], B)
        u, v = dev.structure_to_flow(S, t["epipoles"][f], mu, t["homographies"][f], B)
        assert eq(u.cpu(), u_ref) and eq(v.cpu(), v_ref)
        # paste of combine_flow_simple
        nr = (t["rigidity"] <= 0.5)
        u2, v2 = dev.structure_to_flow(S, t["epipoles"][f], mu, t["homographies"][f], B, nonrigid=cu(nr.astype(np.uint8)),
                                       init_u=cu(t["flow"][f][0]), init_v=cu(t["flow"][f][1]))
        ue, ve = u_ref.copy(), v_ref.copy()
        ue[nr] = t["flow"][f][0][nr]; ve[nr] = t["flow"][f][1][nr]
        assert eq(u2.cpu(), ue) and eq(v2.cpu(), ve)


def test_band_equals_whole(dev, trip):
    """A row band computes bit-identical values to the whole frame (global coordinates)."""
    t = trip
    h, w = t["shape"]
    Hinv = np.linalg.inv(t["homographies"][1])
    whole = dev.flow_to_parallax(cu(t["flow"][1][0]), cu(t["flow"][1][1]), Hinv, t["epipoles"][1])
    y0, y1 = 17, 53
    band = dev.flow_to_parallax(cu(t["flow"][1][0][y0:y1]), cu(t["flow"][1][1][y0:y1]), Hinv, t["epipoles"][1], y0=y0)
    for a, b in zip(whole, band):
        assert eq(a[y0:y1].cpu(), b.cpu())
    s = sum(float(dev.epipole_dist_sum(b - a, w, t["epipoles"][1], y0=a).cpu()[0]) for a, b in ((0, y0), (y0, y1), (y1, h)))
    assert abs(s - float(whole[3].sum().cpu())) <= 1e-12 * s


def test_rigidity_unaries(dev, trip):
    from oracle import hotpath as hp
    t = trip
    h, w = t["shape"]
    y, x = np.mgrid[:h, :w]
    ref = _oracle_front(t)
    for f in range(2):
        q = t["epipoles"][f]
        want = hp.get_unaries_rigid(x, y, ref[f][0], ref[f][1], q[0], q[1], sigma=1.0)
        got = dev.rigidity_unaries(cu(ref[f][0]), cu(ref[f][1]), q, sigma=1.0).cpu().numpy()
        # transcendental functions (atan2, sin, exp, I0e) agree to a few ulp, but the reference
        # evaluates arctan2 on fp32 data (numpy's fp32 loop: 1-4 ulp, CPU dependent), so a few
        # pixels move by an fp32 ulp of the angle: probabilities within 1e-5 relative
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-300)
        assert np.median(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-12


def test_compute_structure_dense_parts(dev, trip):
    """pipeline/compute_structure.py:104-221,374-394 against the oracle's intermediates."""
    from oracle import hotpath as hp
    t = trip
    h, w = t["shape"]
    rs = np.random.RandomState(5)
    occ = [rs.rand(h, w) < 0.05, rs.rand(h, w) < 0.05]
    params = hp.Params()
    out, inter = hp.compute_structure(t["images"], t["flow"], t["rigidity"], occ, t["homographies"], t["epipoles"],
                                      params, infer_mrf=lambda I, un, lam: un[:, :, 1], return_intermediates=True)
    S, _, mu_ar, B_ar, q_new, _, _ = out
    res = inter["residual"]
    o8 = [cu(o.astype(np.uint8)) for o in occ]
    p_dir, p_thr, n_thr = dev.rigidity_direction([cu(res[0][0]), cu(res[0][1])], [cu(res[1][0]), cu(res[1][1])], o8[0], o8[1],
                                          cu(t["rigidity"]), q_new[0], q_new[1], 1.0, params.rigidity_weight_cnn)
    np.testing.assert_allclose(p_dir.cpu().numpy(), inter["p_dir"], rtol=1e-5)
    thr_ref = inter["p_thr"]
    assert int(n_thr.cpu()[0]) == int(p_thr.sum().cpu())
    margin = np.abs(params.rigidity_weight_cnn * t["rigidity"] + (1 - params.rigidity_weight_cnn) * inter["p_dir"] - 0.5)
    assert np.array_equal(p_thr.cpu().numpy()[margin > 1e-5] > 0, thr_ref[margin > 1e-5])

    par = inter["parallax"]
    thr8 = cu(thr_ref.astype(np.uint8))
    pv, cand, cnt = dev.structure_candidates(cu(par[0]), cu(par[1]), thr8, q_new[0], q_new[1], mu_ar[0], mu_ar[1], 1.0, 1)
    assert eq(pv.cpu().numpy() > 0, inter["pv"])
    n = int(cnt.cpu()[0])
    assert n == inter["pv"].sum()
    s1 = hp.parallax2normedstructure(par[1], q_new[1], 1.0, mu_ar[1])[inter["pv"]]
    assert eq(np.sort(cand[:n].cpu().numpy()), np.sort(s1))
    med = dev.median(cand, n)
    assert med == np.median(s1)
    mad = dev.median(dev.abs_deviation(cand, med, n))
    assert 1.426 * mad == B_ar[1]
    _, cand2, cnt2 = dev.structure_candidates(cu(par[0]), cu(par[1]), thr8, q_new[0], q_new[1], mu_ar[0], mu_ar[1],
                                              B_ar[1], 2)
    assert dev.median(cand2, int(cnt2.cpu()[0])) == B_ar[0]

    un, ui = dev.rigidity_structure(cu(S[0]), cu(S[1]), cu(inter["p_dir"]), cu(t["rigidity"]), o8[0], o8[1], mu_ar[0],
                                    mu_ar[1], params.rigidity_sigma_structure, params.rigidity_weight_cnn, want_i32=True)
    np.testing.assert_allclose(un.cpu().numpy(), inter["unaries"], rtol=1e-12, atol=1e-15)
    want_i = ((-inter["unaries"]) * 1000).astype("int32")
    assert np.abs(ui.cpu().numpy() - want_i).max() <= 1      # exp() ulp at an integer boundary


def test_erode(dev):
    import cv2
    rs = np.random.RandomState(2)
    m = (rs.rand(61, 77) > 0.2).astype(np.uint8)
    want = cv2.erode(m, np.ones((3, 3), np.uint8))
    assert eq(dev.erode3x3(cu(m)).cpu(), want)


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4097, 200001])
def test_select_and_median(dev, n):
    rs = np.random.RandomState(n)
    x = rs.standard_normal(n) * 10.0 ** rs.randint(-3, 4, n)
    if n > 4:
        x[rs.randint(0, n, n // 3)] = x[0]            # heavy duplicates
        x[1] = -0.0; x[2] = 0.0; x[3] = np.inf; x[4] = -np.inf
    s = np.sort(x)
    for k in sorted({0, (n - 1) // 2, n // 2, n - 1}):
        r = dev.select_kth(cu(x), k).cpu().numpy()
        assert r[0] == s[k] and r[2] == 0
        if k + 1 < n:
            assert r[1] == s[k + 1]
    assert dev.median(cu(x)) == np.median(x)
    if n > 4:
        x[5 % n] = np.nan
        r = dev.select_kth(cu(x), n - 1).cpu().numpy()
        assert np.isnan(r[0]) and r[2] == 1
        assert np.isnan(dev.median(cu(x)))


def test_masked_minmax(dev, trip):
    from oracle import hotpath as hp
    t = trip
    h, w = t["shape"]
    rs = np.random.RandomState(3)
    S = rs.standard_normal((h, w)) * 3
    rig = (rs.rand(h, w) > 0.6).astype(np.float64)
    q = t["epipoles"][1]
    for mode in ("as_written", "rigid_only"):
        rng = hp.structure_range([S, S], rig, [q, q], range_mask=mode)
        zm = (rig > 0) if mode == "as_written" else ~(rig > 0)
        bx0 = max(0, min(w - 1, int(q[0] - 10))); bx1 = max(0, min(w - 1, int(q[0] + 10)))
        by0 = max(0, min(h - 1, int(q[1] - 10))); by1 = max(0, min(h - 1, int(q[1] + 10)))
        box = None if (bx0 == w - 1 or bx1 == 0 or by0 == h - 1 or by1 == 0) else (bx0, bx1, by0, by1)
        mm = dev.masked_minmax(cu(S), cu(zm.astype(np.uint8)), box).cpu().numpy()
        assert 1.5 * mm[0] == rng[0] and 1.5 * mm[1] == rng[-1]
    S[3, 4] = np.nan
    assert np.isnan(dev.masked_minmax(cu(S)).cpu().numpy()).all()


def test_prepare_channels(dev, trip):
    from oracle import hotpath as hp
    for I in trip["images"]:
        want = hp.prepare_channels(I)
        ch64, tex = dev.prepare_channels(cu(I))
        assert eq(ch64.cpu(), want)
        assert eq(tex.cpu()[..., :3], want.astype(np.float32)) and float(tex[..., 3].abs().max().cpu()) == 0.0
        ch64b, _ = dev.prepare_channels(cu(I.astype(np.float32)), want_texels=False)
        assert eq(ch64b.cpu(), want)


@pytest.mark.parametrize("mode", ["cubic", "linear"])
@pytest.mark.parametrize("cn", [1, 2, 3, 4])
def test_warp_sample_vs_cv2(dev, mode, cn):
    import cv2
    rs = np.random.RandomState(3 + cn)
    h, w = 97, 131
    I = (rs.rand(h, w, cn) * 255).astype(np.float32)
    if cn == 1:
        I = I[..., 0]
    x = (rs.rand(h, w) * (w + 20) - 10)
    y = (rs.rand(h, w) * (h + 20) - 10)
    x[0, :8] = [0, 1, w - 1, w - 2, 5.015625, 5.046875, -0.015625, w - 1 + 0.015625]
    y[0, :8] = [0, h - 1, 0, h - 2, 7.015625, 7.046875, -0.015625, h - 1]
    x[1, :6] = [1e9, -1e9, np.nan, np.inf, -np.inf, 70000.0]
    y[1, :6] = [3.0, 4.0, 5.0, 6.0, 7.0, -70000.0]
    flag = cv2.INTER_CUBIC if mode == "cubic" else cv2.INTER_LINEAR
    want = cv2.remap(I, x.astype(np.float32), y.astype(np.float32), interpolation=flag, borderMode=cv2.BORDER_REPLICATE)
    J, mask = dev.warp_sample(cu(I), None, cu(x), cu(y), interp=mode)
    assert eq(J.cpu(), want)
    with np.errstate(invalid="ignore"):
        m = ((x >= 0) * (x <= w - 1) * (y >= 0) * (y <= h - 1)).astype(float)
    assert eq(mask.cpu(), m)


def test_interpolate_through_H_golden(dev, golden):
    tag, g = golden
    t = golden_triplet(tag)
    J, m = dev.warp_sample(cu(g["ith_I"], None).float(), t["homographies"][1], cu(g["ith_xn"]), cu(g["ith_yn"]))
    assert eq(J.cpu(), g["ith_J"]) and eq(m.cpu(), g["ith_mask"])


# ------------------------------------------------------------------------------------------ a12
def _cv_setup(t, seed=11, with_masks=True):
    from oracle import hotpath as hp
    h, w = t["shape"]
    rs = np.random.RandomState(seed)
    ch = [hp.prepare_channels(I) for I in t["images"]]
    y, x = np.mgrid[:h, :w]
    mu = [np.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2).mean() for q in t["epipoles"]]
    rig = (rs.rand(h, w) > 0.1).astype(np.float64) if with_masks else np.ones((h, w))
    occ = [(rs.rand(h, w) < 0.05) if with_masks else np.zeros((h, w), bool) for _ in range(2)]
    return ch, mu, rig, occ


def _check_costs(got, want, what=""):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    tol = COST_ATOL + COST_RTOL * np.abs(want.astype(np.float64))
    assert (err <= tol).all(), "%s: %d of %d costs outside 1e-5 relative, worst %g" % (
        what, (err > tol).sum(), err.size, (err / np.maximum(np.abs(want), 1e-30)).max())
    return float((got == want).mean())


def _check_argmin(am, cost_ref):
    """argmin must agree wherever the oracle's best-vs-second gap exceeds the cost tolerance."""
    c = np.sort(cost_ref.astype(np.float64), axis=0)
    gap = c[1] - c[0] if c.shape[0] > 1 else np.ones_like(c[0])
    clear = gap > (2 * COST_ATOL + 2 * COST_RTOL * np.abs(c[0]))
    ref = cost_ref.argmin(axis=0)
    assert np.array_equal(np.asarray(am)[clear], ref[clear])
    return float(clear.mean())


@pytest.mark.parametrize("mode", ["cubic", "linear"])
@pytest.mark.parametrize("variant", ["staged", "gather"])
def test_cost_volume_vs_oracle(dev, trip, mode, variant):
    from oracle import hotpath as hp
    t = trip
    ch, mu, rig, occ = _cv_setup(t)
    B = [-0.37, 0.41]
    labels = np.linspace(-3.0, 4.5, 13)
    want, wmin, wam = hp.cost_volume(ch, rig, occ, t["homographies"], B, mu, t["epipoles"], labels,
                                     kind="lorentzian", sigma=1.0, mode=mode)
    nr = cu((rig == 0).astype(np.uint8))
    _, ref_tex = None, None
    ref_ch = cu(ch[1])
    for f, nbi in ((0, 0), (1, 2)):
        tex = cu(np.concatenate([ch[nbi].astype(np.float32), np.zeros(ch[nbi].shape[:2] + (1,), np.float32)], axis=2))
        cost, mn, am = dev.cost_volume(ref_ch, tex, cu(labels), t["homographies"][f], t["epipoles"][f], mu[f], B[f],
                                       label_scale=(mu[0] / mu[1] if f == 0 else 1.0), nonrigid=nr,
                                       occluded=cu(occ[f].astype(np.uint8)), rho="lorentzian", rho_param=1.0,
                                       interp=mode, variant=variant)
        exact = _check_costs(cost.cpu().numpy(), want[f], "frame %d" % f)
        assert exact > 0.999          # log() may differ by an fp64 ulp before the fp32 rounding
        assert eq(mn.cpu(), cost.min(dim=0).values.cpu())
        assert eq(am.cpu(), cost.cpu().numpy().argmin(axis=0))
        _check_argmin(am.cpu().numpy(), want[f])


def test_cost_volume_auto_sigma(dev, trip):
    """The reference's auto-sigma Lorentzian (utils/robust_functions.py:55-58) per label plane: sigma from the
    exact median of |Iz| over the plane.  Against the oracle (same rho with sigma=0 -> auto)."""
    from oracle import hotpath as hp
    t = trip
    ch, mu, rig, occ = _cv_setup(t)
    B = [-0.37, 0.41]
    labels = np.linspace(-3.0, 4.5, 9)
    want, wmin, wam = hp.cost_volume(ch, rig, occ, t["homographies"], B, mu, t["epipoles"], labels, kind="lorentzian", sigma=0.0)
    nr = cu((rig == 0).astype(np.uint8))
    ref_ch = cu(ch[1])
    for f, nbi in ((0, 0), (1, 2)):
        tex = cu(np.concatenate([ch[nbi].astype(np.float32), np.zeros(ch[nbi].shape[:2] + (1,), np.float32)], axis=2))
        cost, mn, am = dev.cost_volume_auto(ref_ch, tex, cu(labels), t["homographies"][f], t["epipoles"][f], mu[f], B[f],
                                            label_scale=(mu[0] / mu[1] if f == 0 else 1.0), nonrigid=nr,
                                            occluded=cu(occ[f].astype(np.uint8)), frame=f)
        exact = _check_costs(cost.cpu().numpy(), want[f], "auto-sigma frame %d" % f)
        assert exact > 0.99
        assert eq(mn.cpu(), cost.min(dim=0).values.cpu()) and eq(am.cpu(), cost.cpu().numpy().argmin(axis=0))
        _check_argmin(am.cpu().numpy(), want[f])


def test_cost_volume_golden_planes(dev, golden):
    """Against planes produced by the reference's own sampler + rho (tests/golden/make_golden.py)."""
    from oracle import hotpath as hp
    tag, g = golden
    t = golden_triplet(tag)
    ch = [hp.prepare_channels(I) for I in t["images"]]
    occ = cu((g["dataterm_occ"] == 1).astype(np.uint8))
    nr = cu((g["dataterm_rig"] == 0).astype(np.uint8))
    mu = list(g["mu"])
    want, _, _ = hp.cost_volume(ch, g["dataterm_rig"], [g["dataterm_occ"]] * 2, t["homographies"], [-0.37, 0.41], mu,
                                t["epipoles"], g["cv_labels"], kind="lorentzian", sigma=1.0)
    got = []
    for f, (B, nbi) in enumerate(((-0.37, 0), (0.41, 2))):
        _, tex = dev.prepare_channels(cu(t["images"][nbi]), want_ch64=False)
        ref_ch, _ = dev.prepare_channels(cu(t["images"][1]), want_texels=False)
        cost, mn, am = dev.cost_volume(ref_ch, tex, cu(g["cv_labels"]), t["homographies"][f], t["epipoles"][f], mu[f], B,
                                       label_scale=(mu[0] / mu[1] if f == 0 else 1.0), nonrigid=nr, occluded=occ)
        got.append(cost.cpu().numpy())
        _check_argmin(am.cpu().numpy(), want[f])
        c = np.sort(want[f].astype(np.float64), axis=0)
        clear = (c[1] - c[0]) > (2 * COST_ATOL + 2 * COST_RTOL * np.abs(c[0]))
        assert np.array_equal(am.cpu().numpy()[clear], g["cv_argmin"][f][clear])
    got = np.stack(got)
    _check_costs(got, want, "oracle planes")
    _check_costs(got[:, :, 20:36, 30:62], g["cv_cost_crop"], "golden crop")
    _check_costs(got.min(axis=1), g["cv_min"], "golden min")


@pytest.mark.parametrize("rho,param", [("charbonnier", 1e-3), ("geman_mcclure", 2.0), ("quadratic", 0.0),
                                       ("lorentzian_auto", 0.7), ("none", 0.0)])
def test_cost_volume_rho_kinds(dev, trip, rho, param):
    from oracle import hotpath as hp
    t = trip
    ch, mu, rig, occ = _cv_setup(t, seed=12)
    labels = np.linspace(-2.0, 3.0, 5)
    f, nbi, B = 1, 2, 0.41
    tex = cu(np.concatenate([ch[nbi].astype(np.float32), np.zeros(ch[nbi].shape[:2] + (1,), np.float32)], axis=2))
    cost, _, _ = dev.cost_volume(cu(ch[1]), tex, cu(labels), t["homographies"][f], t["epipoles"][f], mu[f], B,
                                 nonrigid=cu((rig == 0).astype(np.uint8)), occluded=cu(occ[f].astype(np.uint8)),
                                 rho=rho, rho_param=param)
    planes = []
    for lab in labels:
        Iz = hp.data_residual(lab, ch[1], ch[nbi], rig, occ[f], t["homographies"][f], B, mu[f], t["epipoles"][f])
        x = Iz ** 2
        if rho == "charbonnier":
            v = hp.rho("charbonnier", x, eps=param)
        elif rho == "geman_mcclure":
            v = hp.rho("geman_mcclure", x, sigma=param)
        elif rho == "quadratic":
            v = x
        elif rho == "lorentzian_auto":
            v = param * np.log(1.0 + 0.5 * x / (param ** 2))
        else:
            v = Iz
        planes.append(v.astype(np.float32))
    want = np.stack(planes)
    exact = _check_costs(cost.cpu().numpy(), want, rho)
    if rho in ("charbonnier", "geman_mcclure", "quadratic", "none"):
        assert exact == 1.0          # only IEEE +,-,*,/,sqrt on this route: bit-exact


def test_cost_volume_i32_layout(dev, trip):
    t = trip
    from oracle import hotpath as hp
    ch, mu, rig, occ = _cv_setup(t, seed=13)
    labels = np.linspace(-3.0, 4.5, 9)
    f, nbi, B = 1, 2, 0.41
    tex = cu(np.concatenate([ch[nbi].astype(np.float32), np.zeros(ch[nbi].shape[:2] + (1,), np.float32)], axis=2))
    kw = dict(nonrigid=cu((rig == 0).astype(np.uint8)), occluded=cu(occ[f].astype(np.uint8)))
    cost, _, _ = dev.cost_volume(cu(ch[1]), tex, cu(labels), t["homographies"][f], t["epipoles"][f], mu[f], B, **kw)
    ci, _, am = dev.cost_volume(cu(ch[1]), tex, cu(labels), t["homographies"][f], t["epipoles"][f], mu[f], B,
                                layout="HWL_I32", i32_scale=1000.0, **kw)
    want = (cost.cpu().numpy() * np.float32(1000.0)).astype(np.int32).transpose(1, 2, 0)
    assert ci.shape == want.shape and eq(ci.cpu(), want)


def test_cost_volume_bands_equal_whole(dev, trip):
    """Row bands with a neighbour halo reproduce the whole-frame volume bit for bit; a halo that
    is too small is reported through halo_miss."""
    import torch
    t = trip
    ch, mu, rig, occ = _cv_setup(t, seed=14, with_masks=False)
    h, w = t["shape"]
    labels = np.linspace(-1.0, 2.5, 7)
    f, nbi, B = 1, 2, 0.41
    tex_np = np.concatenate([ch[nbi].astype(np.float32), np.zeros(ch[nbi].shape[:2] + (1,), np.float32)], axis=2)
    args = (t["homographies"][f], t["epipoles"][f], mu[f], B)
    whole, wmn, wam = dev.cost_volume(cu(ch[1]), cu(tex_np), cu(labels), *args)
    halo = 24
    for y0, y1 in ((0, 30), (30, 51), (51, h)):
        n0, n1 = max(0, y0 - halo), min(h, y1 + halo)
        miss = torch.zeros(1, dtype=torch.int32, device="cuda")
        c, mn, am = dev.cost_volume(cu(ch[1][y0:y1]), cu(tex_np[n0:n1]), cu(labels), *args, y0=y0, frame_h=h, nb_y0=n0,
                                    halo_miss=miss)
        assert int(miss.cpu()[0]) == 0
        assert eq(c.cpu(), whole[:, y0:y1].cpu()) and eq(am.cpu(), wam[y0:y1].cpu()) and eq(mn.cpu(), wmn[y0:y1].cpu())
    miss = torch.zeros(1, dtype=torch.int32, device="cuda")
    dev.cost_volume(cu(ch[1][30:40]), cu(tex_np[34:36]), cu(labels), *args, y0=30, frame_h=h, nb_y0=34, halo_miss=miss)
    assert int(miss.cpu()[0]) > 0


@pytest.mark.parametrize("shape,epi", [("sintel", "outside"), ("kitti", "inside")])
def test_full_size_properties(dev, shape, epi):
    """BASELINE.json's full shapes through size-independent properties: staged == gather bit for
    bit, min/argmin consistent with the volume, the planted structure is the argmin label on
    textured pixels, structure -> flow -> structure round trip."""
    import torch
    from mrflow_b200 import synth
    t = synth.make_triplet(shape, seed=1002, epipole=epi, moving_objects=(shape == "kitti"))
    h, w = t["shape"]
    f, nbi = 1, 2
    q, mu, B, H = t["epipoles"][f], t["mu"][f], t["B"][f], t["homographies"][f]
    ref_ch, _ = dev.prepare_channels(cu(t["images"][1]), want_texels=False)
    _, tex = dev.prepare_channels(cu(t["images"][nbi]), want_ch64=False)
    A = t["A_true"]
    labels = np.linspace(1.5 * min(A.min(), 0), 1.5 * A.max(), 51)
    nr = cu((t["rigidity"] <= 0.5).astype(np.uint8))
    c1, mn1, am1 = dev.cost_volume(ref_ch, tex, cu(labels), H, q, mu, B, nonrigid=nr, variant="staged")
    c2, mn2, am2 = dev.cost_volume(ref_ch, tex, cu(labels), H, q, mu, B, nonrigid=nr, variant="gather")
    assert torch.equal(c1, c2) and torch.equal(mn1, mn2) and torch.equal(am1, am2)
    assert torch.equal(mn1, c1.min(dim=0).values)
    # known answer: nearest label to the planted structure, away from borders / moving objects
    step = labels[1] - labels[0]
    planted = np.clip(np.rint((A - labels[0]) / step), 0, 50).astype(np.int64)
    am = am1.cpu().numpy()
    sel = np.zeros((h, w), bool); sel[20:-20, 20:-20] = True
    sel &= t["rigidity"] > 0.5
    d = np.sqrt((np.mgrid[:h, :w][1] - q[0]) ** 2 + (np.mgrid[:h, :w][0] - q[1]) ** 2)
    sel &= d > 60            # parallax vanishes at the epipole: labels are indistinguishable there
    # the planted label is (one of) the best: its cost is ~500x below the average label's, and the
    # argmin lands on it for most pixels (periodic texture gives some pixels a second, equally good
    # label elsewhere -- the CPU oracle has the same 58% on this triplet)
    cost = c1.cpu().numpy()
    at_planted = np.take_along_axis(cost, planted[None], 0)[0]
    assert np.median(at_planted[sel]) < 0.02 * np.median(cost.mean(axis=0)[sel])
    assert (np.abs(am - planted)[sel] <= 1).mean() > 0.45
    # structure -> flow -> structure
    S = cu(A)
    u, v = dev.structure_to_flow(S, q, mu, H, B)
    _, _, par, _ = dev.flow_to_parallax(u, v, np.linalg.inv(H), q)
    S2 = dev.parallax_to_structure(par, q, B, mu).cpu().numpy()
    ok = d > 5
    assert np.abs(S2 - A)[ok].max() < 5e-3


@pytest.mark.parametrize("n", [0, 1, 2, 7, 1000, 50001])
def test_device_side_median_and_label_range(dev, n):
    """The device-resident statistics of the pre-pass: numpy.median with n read on the device, and
    numpy.linspace(1.5*min, 1.5*max, L) from the state's min/max."""
    import ctypes as C
    import torch
    from mrflow_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(100 + n)
    cap = max(n, 1) + 13
    x = rs.standard_normal(cap) * 5
    if n > 4:
        x[rs.randint(0, n, n // 4)] = x[0]
    keys = cu(x)
    cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    scratch = torch.empty(L.mrf_front_scratch_bytes(), dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.mrf_select_median_dev(C.c_void_p(keys.data_ptr()), C.c_void_p(cnt.data_ptr()), cap,
                                       C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()), st), "median_dev")
    got = float(out.cpu()[0])
    if n == 0:
        assert np.isnan(got)
    else:
        assert got == np.median(x[:n])
    state = torch.zeros(_lib.STATE_WORDS, dtype=torch.float64, device="cuda")
    lo0, hi0, lo1, hi1 = -0.37 - n, 2.11, -0.2, 3.3 + 0.001 * n
    state[_lib.STATE_MINMAX:_lib.STATE_MINMAX + 4] = torch.tensor([lo0, hi0, lo1, hi1], dtype=torch.float64)
    for nl in (51, 7, 2):
        _lib.check(L.mrf_label_range_dev(C.c_void_p(state.data_ptr()), nl, None, 0, None, st), "label_range_dev")
        lab = state[_lib.STATE_LABELS:_lib.STATE_LABELS + nl].cpu().numpy()
        assert np.array_equal(lab, np.linspace(1.5 * min(lo0, lo1), 1.5 * max(hi0, hi1), nl))
    state[_lib.STATE_MINMAX:_lib.STATE_MINMAX + 4] = 0.0          # degenerate range: all labels equal
    _lib.check(L.mrf_label_range_dev(C.c_void_p(state.data_ptr()), 51, None, 0, None, st), "label_range_dev")
    assert np.array_equal(state[_lib.STATE_LABELS:_lib.STATE_LABELS + 51].cpu().numpy(), np.linspace(0.0, 0.0, 51))
    # row bands: the six range words of every band are gathered and combined in the kernel
    g = torch.tensor([[lo0, 1.0, 0.5, hi1, 0, 0], [-0.1, hi0, lo1, 2.0, 0, 0], [0.0, 0.0, 0.0, 0.0, 0, 0]], dtype=torch.float64,
                     device="cuda")
    _lib.check(L.mrf_label_range_dev(C.c_void_p(state.data_ptr()), 51, C.c_void_p(g.data_ptr()), 3, None, st), "label_range_dev")
    assert np.array_equal(state[_lib.STATE_LABELS:_lib.STATE_LABELS + 51].cpu().numpy(),
                          np.linspace(1.5 * min(lo0, lo1), 1.5 * max(hi0, hi1), 51))
    assert state[_lib.STATE_MINMAX:_lib.STATE_MINMAX + 4].tolist() == [min(lo0, -0.1, 0.0), max(1.0, hi0), min(0.5, lo1, 0.0), max(hi1, 2.0)]
    g[1, 4] = 1.0                                                  # a NaN flag on one band poisons the label set
    _lib.check(L.mrf_label_range_dev(C.c_void_p(state.data_ptr()), 51, C.c_void_p(g.data_ptr()), 3, None, st), "label_range_dev")
    assert np.isnan(state[_lib.STATE_LABELS:_lib.STATE_LABELS + 51].cpu().numpy()).all()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 257, 4097, 300001, 446464])
@pytest.mark.parametrize("kind", ["normal", "dups", "alldup", "zeros", "nan", "wide", "binade", "binade_big", "inf"])
def test_median_adversarial(dev, n, kind):
    """The exact device-side median against numpy.median on inputs that stress a radix select: one repeated value,
    mostly exact zeros (a masked residual plane), a few thousand or all keys sharing 25 leading bits (every digit is
    needed), wide dynamic range, signed zeros, NaNs, infinities; and the state words the row-band protocol exchanges."""
    import ctypes as C
    import torch
    from mrflow_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(11 * n + len(kind))
    x = rs.standard_normal(n) * (10.0 ** rs.randint(-8, 9, n) if kind == "wide" else 3.0)
    if kind == "dups" and n > 2:
        x[rs.randint(0, n, n // 2)] = 1.25
        x[:2] = [0.0, -0.0]
    if kind == "alldup":
        x[:] = -7.5
    if kind == "zeros" and n > 2:                 # the |Iz| planes of the auto-sigma mode: mostly exact zeros
        x = np.abs(x); x[rs.rand(n) < 0.7] = 0.0
    if kind == "nan" and n > 2:
        x[n // 2] = np.nan
    if kind == "binade" and n > 8:                # the median's class: <= 2000 keys sharing 25 bits -> radix over the list
        d = min(n // 2, 2000)
        lo = (n - d) // 2
        x = np.concatenate([-1.0 - np.abs(x[:lo]), 1.0 + rs.rand(d) * 1e-9, 5.0 + np.abs(x[lo + d:])])
        rs.shuffle(x)
    if kind == "binade_big":                      # ... for all keys: the class cannot be gathered
        x = 1.0 + rs.rand(n) * 1e-9
    if kind == "inf" and n > 3:
        x[0] = np.inf; x[1] = -np.inf; x[2] = np.inf
    cap = n + 7
    keys = cu(np.concatenate([x, np.full(7, 123.0)]))
    cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    scr = torch.zeros(L.mrf_select_scratch_bytes(0), dtype=torch.uint8, device="cuda")
    _lib.check(L.mrf_select_median_dev(P(keys), P(cnt), cap, P(out), P(scr), st), "median_dev")
    got = float(out.cpu()[0])
    with np.errstate(invalid="ignore"):
        want = np.median(x) if n else np.nan
    assert (np.isnan(got) and np.isnan(want)) or got == want, (got, want)
    if n and not np.isnan(want):
        w = scr.view(torch.int64)[:8].cpu().numpy()
        xs = np.sort(x)
        k = (n - 1) // 2
        assert int(w[6]) == n and int(w[3]) == 0 and int(w[4]) == int((x <= xs[k]).sum())


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4097, 300001])
@pytest.mark.parametrize("kind", ["normal", "dups", "nan", "wide"])
def test_select13_sharded_steps(dev, n, kind):
    """The five-pass select in steps over two 'shards' on one GPU -- the calls the row-band pre-pass makes, with
    the NCCL all-reduce / all-gather replaced by the same sums on one device."""
    import ctypes as C
    import torch
    from mrflow_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(7 * n + len(kind))
    x = rs.standard_normal(n) * (10.0 ** rs.randint(-8, 9, n) if kind == "wide" else 3.0)
    if kind == "dups" and n > 2:
        x[rs.randint(0, n, n // 2)] = 1.25
        x[:2] = [0.0, -0.0]
    if kind == "nan" and n > 2:
        x[n // 2] = np.nan
    cut = n // 3
    shards = [x[:cut], x[cut:]]
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nb = L.mrf_select_scratch_bytes(0)
    keys = [cu(np.concatenate([s_, np.zeros(5)])) for s_ in shards]          # capacity > count
    cnt = [torch.tensor([len(s_)], dtype=torch.int32, device="cuda") for s_ in shards]
    scr = [torch.empty(nb, dtype=torch.uint8, device="cuda") for _ in shards]
    o, bins = _lib.SELECT13_HIST_OFFSET // 4, _lib.SELECT13_BINS
    for r in range(2):
        _lib.check(L.mrf_select13_begin(P(scr[r]), st), "begin")
        _lib.check(L.mrf_select13_hist_dev(P(keys[r]), P(cnt[r]), keys[r].numel(), 0, P(scr[r]), None, st), "hist0")
    for ps in range(_lib.SELECT13_PASSES):
        hs = [scr[r].view(torch.int32)[o + ps * bins:o + (ps + 1) * bins] for r in range(2)]
        tot = hs[0] + hs[1]                                               # the all-reduce
        hs[0].copy_(tot); hs[1].copy_(tot)
        for r in range(2):
            if ps + 1 < _lib.SELECT13_PASSES:
                _lib.check(L.mrf_select13_hist_dev(P(keys[r]), P(cnt[r]), keys[r].numel(), ps + 1, P(scr[r]), None, st), "hist")
            else:
                _lib.check(L.mrf_select13_successor_dev(P(keys[r]), P(cnt[r]), keys[r].numel(), P(scr[r]), None, st), "succ")
    so = _lib.SELECT13_SUCC_OFFSET // 8
    gathered = torch.cat([scr[r].view(torch.int64)[so:so + 3] for r in range(2)]).contiguous()   # the all-gather
    for r in range(2):
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib.check(L.mrf_select13_median(P(scr[r]), P(gathered), 2, P(out), st), "median")
        got = float(out.cpu()[0])
        want = np.median(x) if n else np.nan
        assert (np.isnan(got) and np.isnan(want)) or got == want, (n, kind, r, got, want)


@pytest.mark.parametrize("case", ["plain", "few", "none", "nan", "const", "sintel"])
def test_fused_front_equals_stepwise(dev, case):
    """mrf_cost_volume_front as one cooperative kernel against the same pre-pass as separate launches: every
    output (parallax, structures, the device state with mu, B, min/max, labels) bit for bit -- including the
    corners of the exact median: fewer candidates than the gather capacity, none at all, a NaN candidate, and
    candidate ratios spread over many binades."""
    import torch
    from mrflow_b200 import synth
    shape = (436, 1024) if case == "sintel" else (120, 160)
    t = synth.make_triplet(shape, seed=41, epipole="inside" if case == "const" else "outside", flow_noise=0.0 if case == "const" else 0.03,
                           n_waves=4)
    h, w = t["shape"]
    fl = [[cu(t["flow"][f][c].copy()) for c in range(2)] for f in range(2)]
    rigid = np.ones((h, w), np.uint8)
    if case == "few":
        rigid[:] = 0; rigid[40:44, 50:90] = 1
    if case == "none":
        rigid[:] = 0
    if case == "nan":
        fl[1][0][60, 70] = float("nan")
    if case == "const":                     # constant flows, epipole inside the frame: ratios over many binades
        for f in range(2):
            fl[f][0].fill_(1.5 + f); fl[f][1].fill_(-0.75)
    rm = cu(rigid)
    zm = cu((1 - rigid).astype(np.uint8))
    outs = []
    for fused in (True, False):
        buf = dev.FrontBuffers(h, w)
        buf.state.fill_(-7.0)
        dev.cost_volume_front(fl, rm, zm, t["homographies"], t["epipoles"], h, buf, fused=fused)
        torch.cuda.synchronize()
        outs.append(buf)
    a, b = outs
    from mrflow_b200 import _lib
    lab = slice(_lib.STATE_LABELS, _lib.STATE_LABELS + 51)
    sa, sb = a.state.cpu().numpy(), b.state.cpu().numpy()
    assert np.array_equal(sa[:10], sb[:10], equal_nan=True), (sa[:10], sb[:10])
    assert np.array_equal(sa[lab], sb[lab], equal_nan=True)
    for f in range(2):
        assert eq(a.par[f].cpu(), b.par[f].cpu()) and eq(a.S[f].cpu(), b.S[f].cpu())
    n = int(a.count.cpu()[0])
    assert n == int(b.count.cpu()[0])
    if n:
        want = np.median(a.cand[:n].cpu().numpy())
        assert (np.isnan(want) and np.isnan(sa[_lib.STATE_B])) or want == sa[_lib.STATE_B]


def test_resident_pass_equals_host_driven_pass(dev, trip):
    """cost_volume_pass_resident (scalars stay in HBM) against cost_volume_pass (scalars via the host):
    same kernels, so mu, B, labels, structures and both volumes must agree bit for bit."""
    import torch
    from mrflow_b200 import planeparallax as pp
    from mrflow_b200.pipeline import build_cost_volume as bcv
    t = trip
    rig = t["rigidity"].copy(); rig[5:20, 60:80] = 0.0
    band = pp.Band.from_host(t["images"], t["flow"], rig, None)
    mask = cu((rig > 0).astype(np.uint8))
    for mode in ("as_written", "rigid_only"):
        a = bcv.cost_volume_pass(band, t["homographies"], t["epipoles"], mask, pp.LocalStats(), range_mask=mode)
        r = bcv.cost_volume_pass_resident(band, t["homographies"], t["epipoles"], mask, range_mask=mode)
        mu, B, labels = pp.state_to_host(r["buffers"].state)
        assert mu == list(a[2]) and B == list(a[3]) and np.array_equal(labels, a[5])
        for f in range(2):
            assert np.array_equal(r["q"][f], a[4][f])
            assert torch.equal(r["buffers"].S[f], a[1][f]) or eq(r["buffers"].S[f].cpu(), a[1][f].cpu())
            assert torch.equal(r["cost"][f], a[0][f]) and torch.equal(r["argmins"][f], a[7][f])


# ------------------------------------------------------------------------------------------ edge cases
def _tiny_case(h, w, seed):
    rs = np.random.RandomState(seed)
    ch = [rs.rand(h, w, 3) * np.array([255.0, 20.0, 20.0]) for _ in range(3)]
    H = np.array([[1.001, 0.002, 0.4], [-0.001, 0.999, -0.3], [1e-5, -2e-5, 1.0]])
    q = np.array([w * 0.4, h * 0.6])
    y, x = np.mgrid[:h, :w]
    mu = np.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2).mean()
    return ch, H, q, mu


def _run_both(dev, ch, H, q, mu, B, labels, mode="cubic", variant="staged", rig=None, occ=None):
    from oracle import hotpath as hp
    h, w = ch[1].shape[:2]
    rig = np.ones((h, w)) if rig is None else rig
    occ = np.zeros((h, w), bool) if occ is None else occ
    with np.errstate(all="ignore"):
        want, wmin, wam = hp.cost_volume(ch, rig, [occ, occ], [H, H], [B, B], [mu, mu], [q, q], labels, mode=mode, frames=(1,))
    tex = cu(np.concatenate([ch[2].astype(np.float32), np.zeros((h, w, 1), np.float32)], axis=2))
    cost, mn, am = dev.cost_volume(cu(ch[1]), tex, cu(np.asarray(labels, dtype=np.float64)), H, q, mu, B,
                                   nonrigid=cu((rig == 0).astype(np.uint8)), occluded=cu(occ.astype(np.uint8)),
                                   interp=mode, variant=variant)
    return cost.cpu().numpy(), mn.cpu().numpy(), am.cpu().numpy(), want[0], wmin[0], wam[0]


@pytest.mark.parametrize("shape", [(2, 2), (3, 5), (5, 7), (9, 33), (8, 32), (17, 31)])
@pytest.mark.parametrize("mode", ["cubic", "linear"])
def test_cost_volume_tiny_and_ragged_frames(dev, shape, mode):
    """Frames smaller than a tile, smaller than the bicubic footprint, and with ragged tile edges: every
    sample is a border sample (OpenCV's tap-by-tap order, replicated taps)."""
    ch, H, q, mu = _tiny_case(*shape, seed=shape[0] * 100 + shape[1])
    labels = np.linspace(-2.0, 3.0, 9)
    for variant in ("staged", "gather"):
        got, mn, am, want, wmin, wam = _run_both(dev, ch, H, q, mu, 0.6, labels, mode=mode, variant=variant)
        _check_costs(got, want, "%s %s" % (shape, variant))
        _check_argmin(am, want)
        assert eq(mn, got.min(axis=0))


def test_cost_volume_label_extremes(dev):
    """One label, 256 labels, a label exactly at the pole of the parallax formula (B*A/mu == 1: division
    by zero in the reference), huge / infinite / NaN labels: the cold exact path must reproduce numpy's
    inf / NaN propagation (samples leave the frame -> residual 0 -> rho(0))."""
    ch, H, q, mu = _tiny_case(24, 40, seed=5)
    B = 0.5
    pole = mu / B                                    # B * a / mu - 1 == 0 exactly for this a? check both sides too
    for labels in ([1.3], list(np.linspace(-4, 6, 256)), [0.5, pole, np.nextafter(pole, 0), np.nextafter(pole, 1e9), 2.0],
                   [1e300, -1e300, np.inf, -np.inf, np.nan, 1.0], [0.0, -0.0, 1e-310, 1.0]):
        got, mn, am, want, wmin, wam = _run_both(dev, ch, H, q, mu, B, labels)
        assert not np.isnan(got).any()
        _check_costs(got, want, "labels %s.." % str(labels[:3]))
        _check_argmin(am, want)


def test_cost_volume_degenerate_inputs(dev):
    """Non-finite flow-free inputs: NaN / inf in the homography, epipole on a pixel centre (|v| = 0, the 0.5
    clamp of :212), mu tiny, all pixels non-rigid or occluded."""
    ch, H, q, mu = _tiny_case(16, 48, seed=6)
    labels = np.linspace(-1.0, 2.0, 5)
    # epipole exactly on a pixel
    got, _, am, want, _, _ = _run_both(dev, ch, H, np.array([20.0, 7.0]), mu, 0.4, labels)
    _check_costs(got, want, "epipole on pixel"); _check_argmin(am, want)
    # projective row that sends a line of the frame to infinity (den crosses 0 inside the frame)
    Hs = H.copy(); Hs[2] = [0.05, 0.0, -1.0]
    got, _, am, want, _, _ = _run_both(dev, ch, Hs, q, mu, 0.4, labels)
    _check_costs(got, want, "singular den"); _check_argmin(am, want)
    for bad in (np.nan, np.inf):
        Hb = H.copy(); Hb[0, 2] = bad
        got, _, _, want, _, _ = _run_both(dev, ch, Hb, q, mu, 0.4, labels)
        _check_costs(got, want, "H with %s" % bad)
    # everything masked
    h, w = ch[1].shape[:2]
    got, mn, am, want, _, _ = _run_both(dev, ch, H, q, mu, 0.4, labels, rig=np.zeros((h, w)))
    assert (got == 0).all() and (am == 0).all()
    got, _, _, want, _, _ = _run_both(dev, ch, H, q, mu, 0.4, labels, occ=np.ones((h, w), bool))
    assert (got == 0).all()
