"""``pipeline/compute_structure.py`` of the reference on the B200.

``compute_structure(images, flow, rigidity, occlusions_precomputed, homographies, epipoles, params)``
keeps the reference signature (:37) and its 7-element return (:443-449); every dense per-pixel step
runs as a CUDA kernel, the three medians as an exact radix select.  What stays on the host is what
the reference itself delegates to host solvers: the 441-pixel BFGS of ``correct_epipole`` (:453-511),
``refine_A`` (chumpy, :259-356) and the MRF inference of ``rigidity/inference.py:99-135`` (C++ TRW-S /
alpha-expansion in extern/eff_trws) -- the unchanged consumer of the unaries.
"""
import numpy as np

from .. import device, planeparallax as pp
from .._host import down, flag, require_cuda, up
from .build_cost_volume import _down_many


def _default_infer_mrf():
    try:
        from rigidity import inference          # the reference's own module, when it is on sys.path
        return inference.infer_mrf
    except Exception as e:                      # noqa: BLE001
        raise RuntimeError("rigidity.inference.infer_mrf (reference, needs extern/eff_trws) is not importable; "
                           "pass infer_mrf= explicitly") from e


def _refine_A():
    try:
        from initialization import refine_A     # reference module; needs chumpy
        return refine_A.refine_A
    except Exception:                           # noqa: BLE001
        return None


def parallax2normedstructure(parallax, q, B, mu):
    """pipeline/compute_structure.py:25-33"""
    require_cuda()
    return down(device.parallax_to_structure(up(parallax, np.float64), q, B, mu))


def flow2parallax(u, v, q):
    """pipeline/compute_structure.py:517-537.  ``u, v``: residual flow (fp32 as get_residual_flow returns
    it, or fp64).  Returns (parallax, foe_vectors, foe_dists) as fp64 arrays."""
    require_cuda()
    par, vec, dst = device.flow2parallax(up(u, np.float64), up(v, np.float64), q)
    return down(par), down(vec), down(dst)


def correct_epipole(u, v, q, rigidity):
    """pipeline/compute_structure.py:453-511 (host BFGS on the 21x21 window round the epipole)."""
    import torch
    return pp.correct_epipole(torch.from_numpy(np.ascontiguousarray(u)), torch.from_numpy(np.ascontiguousarray(v)), q,
                              torch.from_numpy(np.ascontiguousarray(np.asarray(rigidity).astype(np.uint8))))


def _occ_any(band, occlusions):
    """The structure-consistency step treats ANY non-zero occlusion value as occluded
    (``((occ[0]==0)*(occ[1]==0))==0``, :383) while the direction unaries test ``== 1`` (:130-147).  For boolean
    maps the two agree and the band's masks are reused; otherwise the ``!= 0`` masks are uploaded."""
    arrs = [np.asarray(o) for o in occlusions]
    if all(a.dtype == np.bool_ for a in arrs):
        return band.occ
    return [flag(a != 0) for a in arrs]


def compute_structure(images, flow, rigidity, occlusions_precomputed, homographies, epipoles, params,
                      infer_mrf=None, stats=None):
    """pipeline/compute_structure.py:37-449.

    Returns ``[structure_array, homographies, mu_array, B_array, epipoles_new_array,
    rigidity_refined, occ]`` with the reference's types (structures fp64 [h,w] numpy arrays).
    ``infer_mrf(I, unaries, lambd)`` defaults to the reference's ``rigidity.inference.infer_mrf``.
    Raises ``Exception('TooMuchNonRigid')`` / ``Exception('TooFewStructureMatches')`` like the
    reference (:165, :429) so that mrflow.py's fallback (:153-161) keeps working.
    """
    require_cuda()
    stats = stats or pp.LocalStats()
    rigidity = np.asarray(rigidity)
    h, w = rigidity.shape
    import torch
    band = pp.Band.from_host(None, flow, rigidity, occlusions_precomputed)    # the frames themselves are not sampled here

    if not pp.param(params, "debug_use_rigidity_gt"):
        rigid_mask = device.erode3x3((band.rigidity > 0.5).to(torch.uint8))   # :44
    else:
        rigid_mask = (band.rigidity > 0).to(torch.uint8)                      # :47 (used as `> 0`, :77)

    wc = pp.param(params, "rigidity_weight_cnn")
    scale_structure = pp.param(params, "scale_structure")
    resident = isinstance(stats, pp.LocalStats) and scale_structure != 2
    if resident:
        # Device-resident stage: mu, the valid mask, B_fwd (MAD) and B_b (median ratio) never leave HBM
        # (device.structure_scales_dev); the host is synchronised ONCE, when the results are downloaded -- unless
        # an epipole lies inside the frame (441-pixel BFGS, :453-511) or refine_A is available (chumpy).
        q_new = pp.refine_epipoles(band, homographies, epipoles, rigid_mask, stats)                       # :77
        res, par = [], []
        for f in range(2):
            Hinv = np.linalg.inv(np.asarray(homographies[f], dtype=np.float64))
            ur, vr, p, _ = device.flow_to_parallax(band.flow[f][0], band.flow[f][1], Hinv, q_new[f], want_dists=False)   # :74,:82
            res.append((ur, vr)); par.append(p)
        p_dir, p_thr, n_thr = device.rigidity_direction(res[0], res[1], band.occ[0], band.occ[1], band.rigidity, q_new[0],
                                                        q_new[1], pp.param(params, "rigidity_sigma_direction"), wc)   # :104-153
        sbuf = device.StructureBuffers(h, w, band.device)
        device.structure_scales_dev(par[0], par[1], p_thr, q_new[0], q_new[1], h, sbuf, scale_structure)   # :84,:172-223
        S = [device.parallax_to_structure_dev(par[f], q_new[f], sbuf.state, f) for f in range(2)]         # :240-246
        refine = _refine_A() if pp.param(params, "nonlinear_structure_initialization") > 0 else None
        un = None
        if refine is None and pp.param(params, "rigidity_use_structure"):
            if pp.param(params, "occlusion_reasoning") > 0:
                occ_dev = _occ_any(band, occlusions_precomputed)
            else:
                occ_dev = [flag(np.ones((h, w)) > 0), flag(np.ones((h, w)) > 0)]
            un, _ = device.rigidity_structure_dev(S[0], S[1], p_dir, band.rigidity, occ_dev[0], occ_dev[1], sbuf.state,
                                                  pp.param(params, "rigidity_sigma_structure"), wc)       # :374-394
        small = torch.cat((sbuf.state[:4], n_thr.to(torch.float64)))
        hosts = _down_many([S[0], S[1], small] + ([un] if un is not None else []))                        # the one synchronisation
        S_host = [hosts[0], hosts[1]]
        mu_ar = [float(hosts[2][0]), float(hosts[2][1])]
        B_ar = [float(hosts[2][2]), float(hosts[2][3])]
        if int(hosts[2][4]) < h * w * 0.25:
            raise Exception("TooMuchNonRigid")                                # :164-165
        un_host = hosts[3] if un is not None else None
    else:
        fp = pp.first_pass(band, homographies, epipoles, rigid_mask, stats)    # :67-91
        q_new, mu_ar = fp["q"], fp["mu"]
        p_dir, p_thr, n_thr = device.rigidity_direction(fp["residual"][0], fp["residual"][1], band.occ[0], band.occ[1],
                                                        band.rigidity, q_new[0], q_new[1],
                                                        pp.param(params, "rigidity_sigma_direction"), wc)   # :104-153
        if int(n_thr.cpu()[0]) < h * w * 0.25:
            raise Exception("TooMuchNonRigid")                                # :164-165
        pv, B_ar = pp.scales(band, fp, p_thr, stats, scale_structure)          # :172-223
        S = pp.structures(band, fp, B_ar)                                      # :240-246
        S_host = [down(S[0]), down(S[1])]
        un_host = None
    homographies = list(homographies) if not isinstance(homographies, list) else homographies

    if pp.param(params, "nonlinear_structure_initialization") > 0:            # :259-356
        refine = _refine_A()
        if refine is None:
            print("(WW) initialization.refine_A (needs chumpy) is not importable: backward H/B/q are not refined")
        else:
            occ_ok = (np.asarray(occlusions_precomputed[0]) == 0) * (np.asarray(occlusions_precomputed[1]) == 0)
            A_new, H_new, B_new, q_b = refine(flow[0][0], flow[0][1], homographies[0], B_ar[0], q_new[0], mu_ar[0],
                                              mu_ar[1], S_host[1], down(p_thr) > 0, occ_ok)
            S_host[0] = A_new
            S[0] = up(A_new, np.float64)
            homographies[0] = H_new                                           # :349 (in-place update)
            B_ar[0] = B_new
            q_new[0] = q_b

    if pp.param(params, "occlusion_reasoning") > 0:                           # :364-367
        occ = occlusions_precomputed
        occ_dev = _occ_any(band, occlusions_precomputed)
    else:
        occ = [np.ones((h, w)) > 0, np.ones((h, w)) > 0]
        occ_dev = [flag(occ[0]), flag(occ[1])]

    rigidity_refined = None
    if pp.param(params, "rigidity_use_structure"):                            # :374-398
        if un_host is None:
            un, _ = device.rigidity_structure(S[0], S[1], p_dir, band.rigidity, occ_dev[0], occ_dev[1], mu_ar[0], mu_ar[1],
                                              pp.param(params, "rigidity_sigma_structure"), wc)
            un_host = down(un)
        infer = infer_mrf or _default_infer_mrf()
        rigidity_refined = infer(np.asarray(images[1]), un_host, 1.1) > 0.5   # :396-398
    if pp.param(params, "debug_use_rigidity_gt"):
        rigidity_refined = rigidity                                           # :424-425
    if rigidity_refined is None:
        raise UnboundLocalError("rigidity_refined (rigidity_use_structure=0 without debug_use_rigidity_gt, as in the reference)")
    if (rigidity_refined == 1).sum() < 0.25 * rigidity_refined.size:
        raise Exception("TooFewStructureMatches")                             # :427-429

    return [S_host, homographies, mu_ar, B_ar, q_new, rigidity_refined, occ]
