"""Cost-volume parity at the BASELINE.json shapes (the sizes bench.py times), against the CPU oracle.

The small fixtures of test_gpu_kernels.py never trigger the clipping of the staged window (kWinMaxW, the
48 KB capacity, the band edge); these frames do.  Sintel 1024x436 with the epipole outside the frame,
KITTI 1242x375 with the epipole inside and moving objects, and one row band (with halo) of a 4K frame.
The volume is built over all 51 labels by the whole pipeline (pre-pass -> labels -> cost volume), staged
and gather; the oracle is evaluated on a subset of the label planes (both ends of the range included) so
the CPU side stays at a few seconds.  Bars as everywhere: fp32 costs within 1e-5 relative, argmin exact
wherever the best-vs-second gap exceeds the tolerance.  Template: optimize_structure_single_scale.py:210-302;
label rule pipeline/build_cost_volume.py:190-192.
"""
import numpy as np
import pytest

from test_gpu_kernels import COST_ATOL, COST_RTOL, _check_argmin, _check_costs, cu

pytestmark = pytest.mark.gpu

PLANES = [0, 7, 19, 25, 33, 44, 50]       # both ends of the label range included


def _volumes(t, variant, range_mask="rigid_only"):
    import torch
    from mrflow_b200 import planeparallax as pp
    from mrflow_b200.pipeline.build_cost_volume import build_cost_volume
    out = build_cost_volume(t["images"], t["flow"], t["rigidity"], t["homographies"], t["epipoles"], None,
                            range_mask=range_mask, variant=variant, return_device=True)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("shape,epi,moving", [("sintel", "outside", False), ("kitti", "inside", True)])
def test_pipeline_volume_vs_oracle_at_bench_shapes(shape, epi, moving):
    import torch
    from mrflow_b200 import synth
    from oracle import hotpath as hp
    t = synth.make_triplet(shape, seed=1001 if shape == "sintel" else 1003, epipole=epi, moving_objects=moving)
    h, w = t["shape"]
    rig = t["rigidity"] if not moving else np.where(t["rigidity"] > 0.5, t["rigidity"], 0.0)
    t = dict(t, rigidity=rig)
    st = _volumes(t, "staged")
    ga = _volumes(t, "gather")
    cost_s, S, mu, B, q, labels, (mins, argmins) = st
    for f in range(2):                                   # the two variants are the same arithmetic
        assert torch.equal(cost_s[f], ga[0][f]) and torch.equal(argmins[f], ga[6][1][f])
    assert labels.shape == (51,) and np.isfinite(labels).all() and labels[0] < labels[-1]

    # the oracle's own pre-pass: mu, B_b, labels agree to reduction-order noise
    x, y = np.meshgrid(np.arange(w), np.arange(h))
    par, q_o, mu_o, dn = [], [], [], []
    for f in range(2):
        ur, vr = hp.get_residual_flow(x, y, t["flow"][f][0], t["flow"][f][1], t["homographies"][f])
        qn = hp.correct_epipole(ur, vr, t["epipoles"][f], rig > 0)
        p, _, d = hp.flow2parallax(ur, vr, qn)
        par.append(p); q_o.append(qn); mu_o.append(d.mean()); dn.append(d / mu_o[-1])
    pv = (np.abs(par[0]) > 0.1) & (np.abs(par[1]) > 0.1) & (rig > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        target = par[1][pv] / (1.0 * (par[1][pv] / mu_o[1] - dn[1][pv]))
        template = mu_o[1] / mu_o[0] * par[0][pv] / (par[0][pv] / mu_o[0] - dn[0][pv])
        B_o = [np.median(template / target), 1.0]
    S_o = [hp.parallax2normedstructure(par[f], q_o[f], B_o[f], mu_o[f]) for f in range(2)]
    labels_o = hp.structure_range(S_o, rig, q_o, range_mask="rigid_only")
    np.testing.assert_allclose(np.asarray(q), np.asarray(q_o), rtol=0, atol=1e-9)
    np.testing.assert_allclose(mu, mu_o, rtol=1e-13)
    np.testing.assert_allclose(B, B_o, rtol=1e-10)
    np.testing.assert_allclose(labels, labels_o, rtol=1e-9, atol=1e-12)

    # label planes: oracle evaluated at the GPU's own scalars (bit-exact inputs)
    ch = [hp.prepare_channels(I) for I in t["images"]]
    noocc = [np.zeros((h, w), bool)] * 2
    want, _, _ = hp.cost_volume(ch, rig, noocc, t["homographies"], B, mu, q, labels, labels=PLANES)
    exact = []
    for f in range(2):
        got = cost_s[f][PLANES].cpu().numpy()
        exact.append(_check_costs(got, want[f], "%s frame %d" % (shape, f)))
    assert min(exact) > 0.999, exact
    # ... and at the oracle's own scalars: all but a vanishing fraction of the samples agree
    want_o, _, _ = hp.cost_volume(ch, rig, noocc, t["homographies"], B_o, mu_o, q_o, labels_o, labels=PLANES, frames=(1,))
    close = np.isclose(cost_s[1][PLANES].cpu().numpy(), want_o[0], rtol=COST_RTOL, atol=COST_ATOL)
    print("%s: bit-exact fraction at GPU scalars %s, within 1e-5 at oracle scalars %.6f" % (shape, exact, close.mean()))
    assert close.mean() > 0.999

    # argmin over all 51 labels of the forward frame
    full, _, _ = hp.cost_volume(ch, rig, noocc, t["homographies"], B, mu, q, labels, frames=(1,))
    clear = _check_argmin(argmins[1].cpu().numpy(), full[0])
    assert clear > 0.5
    assert torch.equal(mins[1], cost_s[1].min(dim=0).values)


def test_4k_band_with_halo_vs_oracle():
    """Rows 1080..1352 of a 4K frame with a neighbour halo: two label planes against the oracle run on
    the whole frame, and the band equal to the same rows of the staged whole-frame volume."""
    import torch
    from mrflow_b200 import device as dev, sharding, synth
    from oracle import hotpath as hp
    td = synth.make_triplet_device("4k", seed=1001, n_waves=6)
    h, w = td["shape"]
    f, nbi = 1, 2
    H, q, mu, B = td["homographies"][f], td["epipoles"][f], td["mu"][f], td["B"][f]
    A = td["A_true"]
    labels = np.linspace(1.5 * min(float(A.min()), 0.0), 1.5 * float(A.max()), 51)
    y0, y1 = 1080, 1352
    up, dn = sharding.required_halo((y0, y1), w, h, H, q, mu, B, labels)
    n0, n1 = max(0, y0 - up), min(h, y1 + dn)
    ref_ch, _ = dev.prepare_channels(td["images"][1], want_texels=False)
    _, tex = dev.prepare_channels(td["images"][nbi], want_ch64=False)
    miss = torch.zeros(1, dtype=torch.int32, device="cuda")
    labels_d = cu(labels)
    planes = [6, 41]
    band, mn, am = dev.cost_volume(ref_ch[y0:y1].contiguous(), tex[n0:n1].contiguous(), labels_d, H, q, mu, B, y0=y0,
                                   frame_h=h, nb_y0=n0, halo_miss=miss)
    assert int(miss.cpu()[0]) == 0
    whole, wmn, wam = dev.cost_volume(ref_ch, tex, labels_d, H, q, mu, B, want_cost=False)
    assert torch.equal(am, wam[y0:y1]) and torch.equal(mn, wmn[y0:y1])
    ch1 = hp.prepare_channels(td["images"][1].cpu().numpy())
    ch2 = hp.prepare_channels(td["images"][nbi].cpu().numpy())
    np.testing.assert_array_equal(ref_ch.cpu().numpy(), ch1)
    rig = np.ones((h, w)); occ = np.zeros((h, w), bool)
    for l in planes:
        Iz = hp.data_residual(labels[l], ch1, ch2, rig, occ, H, B, mu, q)
        want = hp.rho("lorentzian", Iz ** 2, sigma=1.0).astype(np.float32)
        assert _check_costs(band[l].cpu().numpy(), want[y0:y1], "4K band plane %d" % l) > 0.999
