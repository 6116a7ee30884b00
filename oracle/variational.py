"""numpy/scipy restatement of the variational structure refinement -- TEST INFRASTRUCTURE ONLY.

structure_refinement/optimize_structure_single_scale.py:310-652 (optimizeStructureSingleScale): IRLS outer
loop around the linearised data term (computeDataTerm, oracle.hotpath.compute_data_term), image-adaptive
first- and second-order robust smoothness, an optional consistency term, a MINRES solve, clipping of the
update and a 7x7 median.  Same third-party calls as the reference (scipy.sparse algebra,
scipy.sparse.linalg.minres, scipy.ndimage.median_filter, cv2.blur / erode / cvtColor); the finite-difference
matrices of structure_refinement/construct_matrices.py:14-38 (utils/matrix_from_filter.py, replicated
borders) are written as explicit stencils and checked against the reference's matrices by
tests/golden/make_golden.py.
"""
import numpy as np
import cv2
from scipy import ndimage, sparse
from scipy.sparse import linalg as spla

from . import hotpath as hp


# --------------------------------------------------------------------------- construct_matrices.py
def gradient_matrix(shape):
    """construct_gradient_matrix (:14-24): vstack(Dx, Dy), correlation with [0,-1,1] and its transpose,
    last valid value replicated at the border (so the last column / row of the derivative is 0)."""
    h, w = shape
    n = h * w
    idx = np.arange(n).reshape(h, w)
    def fwd(axis):
        nxt = np.minimum(np.indices((h, w))[axis] + 1, (h, w)[axis] - 1)
        j = idx[nxt, np.indices((h, w))[1]] if axis == 0 else idx[np.indices((h, w))[0], nxt]
        M = sparse.coo_matrix((np.ones(n), (idx.ravel(), j.ravel())), shape=(n, n)) - sparse.identity(n)
        return M.tocsr()
    return sparse.vstack((fwd(1), fwd(0))).tocsr()


def second_order_matrix(shape):
    """construct_2ndorder_matrix (:26-38): vstack(Dxx, Dyy), correlation with [1,-2,1], replicated border."""
    h, w = shape
    n = h * w
    idx = np.arange(n).reshape(h, w)
    yy, xx = np.indices((h, w))
    def sec(axis):
        if axis == 1:
            a, b = idx[yy, np.maximum(xx - 1, 0)], idx[yy, np.minimum(xx + 1, w - 1)]
        else:
            a, b = idx[np.maximum(yy - 1, 0), xx], idx[np.minimum(yy + 1, h - 1), xx]
        M = (sparse.coo_matrix((np.ones(n), (idx.ravel(), a.ravel())), shape=(n, n))
             + sparse.coo_matrix((np.ones(n), (idx.ravel(), b.ravel())), shape=(n, n)) - 2.0 * sparse.identity(n))
        return M.tocsr()
    return sparse.vstack((sec(1), sec(0))).tocsr()


# --------------------------------------------------------------------------- utils/edge_weights.py
def weights_locally_normalized(I, norm_radius=45):
    """utils/edge_weights.py:38-55 (computeWeightsLocallyNormalized, centered gradient)."""
    gy, gx = np.gradient(I)[:2]
    gysq = (gy ** 2).mean(axis=2) if gy.ndim > 2 else gy ** 2
    gxsq = (gx ** 2).mean(axis=2) if gx.ndim > 2 else gx ** 2
    gx_mean = cv2.blur(gxsq, ksize=(norm_radius, norm_radius))
    gy_mean = cv2.blur(gysq, ksize=(norm_radius, norm_radius))
    return (np.exp(-gxsq * 1.0 / (2 * np.maximum(1e-6, gx_mean))), np.exp(-gysq * 1.0 / (2 * np.maximum(1e-6, gy_mean))))


def clip_dA(dA, A, B, mu, q, threshold=0.5):
    """optimize_structure_single_scale.py:157-183."""
    h, w = A.shape[:2]
    y, x = np.mgrid[:h, :w]
    vx = q[0] - x
    vy = q[1] - y
    n = np.sqrt(vx ** 2 + vy ** 2) / mu
    with np.errstate(divide="ignore", invalid="ignore"):
        num = -(mu ** 2 * threshold - 2 * mu * threshold * A * B + threshold * A ** 2 * B ** 2)
        dA_max = num / (threshold * A * B ** 2 + (-mu ** 2 * n - mu * threshold) * B)
        dA_min = num / (threshold * A * B ** 2 + (mu ** 2 * n - mu * threshold) * B)
    return np.maximum(dA_min, np.minimum(dA.reshape((h, w)), dA_max))


def _robust(kind):
    return (lambda x: hp.rho(kind, x)), (lambda x: hp.drho(kind, x))     # auto-sigma forms (sigma = eps = 0)


def optimize_structure_single_scale(A_init, A_bwd_reference, A_fwd_reference, occlusions_bwd, occlusions_fwd,
                                    occlusions_both, rigidity, images, H_ar, B_ar, q_ar, mu_ar, lambda_1storder,
                                    lambda_2ndorder, lambda_consistency, A_scale_factor=1.0, N_outer=10, N_inner=1,
                                    min_weight=0.0, return_trace=False, weights_ulp_jitter=None):
    """optimize_structure_single_scale.py:310-652.  ``images`` are the three RGB frames (fp64 [h,w,3] in [0,1]).
    Returns (A, energies) -- and, with return_trace, a list of per-iteration dicts for the kernel tests.

    ``weights_ulp_jitter`` (a seed) moves every fp32 edge weight by one unit in the last place, up or down at
    random: the tests use it to measure how far the reference's own result moves under the rounding freedom of
    its fp32 ``np.exp`` (MINRES at rtol 1e-5 on this stiff system amplifies such changes), which is the yardstick
    for the GPU path whose expf rounds differently."""
    robust_ev, robust_deriv = _robust("lorentzian")                       # :333
    robust_ev_c, robust_deriv_c = _robust("charbonnier")                  # :336
    afac_sq = A_scale_factor ** 2
    ch = [hp.prepare_channels(I) for I in images]                          # :350-371
    cw = hp.CHANNEL_WEIGHTS
    h, w = ch[1].shape[:2]
    weights_x, weights_y = weights_locally_normalized(
        cv2.cvtColor(images[1].astype("float32"), cv2.COLOR_RGB2GRAY))    # :381-383
    if weights_ulp_jitter is not None:
        rs = np.random.RandomState(weights_ulp_jitter)
        for wgt in (weights_x, weights_y):
            up_ = rs.rand(*wgt.shape) < 0.5
            wgt[...] = np.where(up_, np.nextafter(wgt, np.float32(2)), np.nextafter(wgt, np.float32(-1)))
    rigidity_eroded = cv2.erode(rigidity.astype("uint8"), kernel=None)    # :387
    weights_x[rigidity_eroded == 0] = 0
    weights_y[rigidity_eroded == 0] = 0
    weights_x = np.maximum(weights_x, min_weight)
    weights_y = np.maximum(weights_y, min_weight)
    A = A_init.copy()
    energies, trace = [], []
    M_GRAD = gradient_matrix((h, w))
    M_DIV = -M_GRAD.T
    M_GRAD2 = second_order_matrix((h, w))
    M_DIV2 = M_GRAD2.T
    for k in range(N_outer):
        diag_f, b_f, E_f = hp.compute_data_term(A, ch[1], ch[2], rigidity, occlusions_fwd, H_ar[1], B_ar[1], mu_ar[1],
                                                q_ar[1], cw=cw)            # :425-432
        diag_b, b_b, E_b = hp.compute_data_term(A * (mu_ar[0] / mu_ar[1]), ch[1], ch[0], rigidity, occlusions_bwd, H_ar[0],
                                                B_ar[0], mu_ar[0], q_ar[0], cw=cw)   # :435-442
        diag_b = diag_b * (mu_ar[0] / mu_ar[1])                            # :448-449
        b_b = b_b * (mu_ar[0] / mu_ar[1])
        D_fwd = diag_f.ravel().copy()
        b_fwd = b_f.ravel().copy()
        rep = occlusions_fwd.ravel() == 1                                  # :452-461
        D_fwd[rep] = diag_b.ravel()[rep]
        b_fwd[rep] = b_b.ravel()[rep]
        # consistency term :476-497
        A_bwd_ref = A_bwd_reference * (mu_ar[1] / mu_ar[0])
        A_fwd_ref = A_fwd_reference
        A_ref_max = np.maximum(A_fwd_ref, A_bwd_ref)
        A_ref_min = np.minimum(A_fwd_ref, A_bwd_ref)
        Iz = (A - A_ref_max) * (A > A_ref_max) + (A - A_ref_min) * (A < A_ref_min)
        Iz[occlusions_bwd == 1] = (A - A_fwd_ref)[occlusions_bwd == 1]
        Iz[occlusions_fwd == 1] = (A - A_bwd_ref)[occlusions_fwd == 1]
        psi_c = afac_sq * robust_deriv_c(afac_sq * Iz ** 2)
        A_cons = sparse.diags(psi_c.flatten(), shape=(h * w, h * w))
        b_cons = A_cons.dot(Iz.flatten())
        E_cons = robust_ev_c(afac_sq * Iz ** 2).sum()
        # 1st order :505-516
        g1 = M_GRAD.dot(A.flatten()).reshape((-1, h * w))
        wa1 = np.c_[weights_x.flatten(), weights_y.flatten()].T
        arg1 = afac_sq * (wa1 * g1 ** 2).sum(axis=0)
        psi1 = afac_sq * robust_deriv(arg1)
        P = sparse.diags(np.r_[(weights_x.flatten() * psi1).flatten(), (weights_y.flatten() * psi1).flatten()])
        A_s1 = -M_DIV.dot(P).dot(M_GRAD)
        b_s1 = A_s1 * A.flatten()
        # 2nd order :526-550
        weights_x[rigidity_eroded == 0] = min_weight
        weights_y[rigidity_eroded == 0] = min_weight
        g2 = M_GRAD2.dot(A.flatten()).reshape((-1, h * w))
        wa2 = np.c_[weights_x.flatten(), weights_y.flatten()].T
        arg2 = afac_sq * (wa2 * g2 ** 2).sum(axis=0)
        psi2 = afac_sq * robust_deriv(arg2)
        P2 = sparse.diags(np.r_[psi2.flatten() * weights_x.flatten(), psi2.flatten() * weights_y.flatten()])
        A_s2 = M_DIV2.dot(P2).dot(M_GRAD2)
        b_s2 = A_s2 * A.flatten()
        E1 = robust_ev(arg1).sum()                                         # :560-561
        E2 = robust_ev(arg2).sum()
        energy = E_f + E_b + lambda_1storder * E1 + lambda_2ndorder * E2 + lambda_consistency * E_cons   # :564
        energies.append(energy / (h * w))
        A_comb = (sparse.diags(D_fwd, shape=(h * w, h * w)) + lambda_1storder * A_s1 + lambda_2ndorder * A_s2
                  + lambda_consistency * A_cons)                            # :580 (A_data_bwd was emptied, :460)
        b_comb = -b_fwd - lambda_1storder * b_s1 - lambda_2ndorder * b_s2 - lambda_consistency * b_cons   # :581
        z = sparse.diags((rigidity == 1).ravel().astype("float32"), shape=(h * w, h * w))   # :586-588
        A_comb = z.dot(A_comb).dot(z)
        b_comb = z.dot(b_comb)
        A_comb = A_comb + sparse.diags(np.full(h * w, 0.01)).tocsr()        # :591-595
        da = spla.minres(A_comb, b_comb, rtol=1e-5)[0]                      # :593  solve(..., 'minres', tol=1e-5)
        da[np.isnan(da)] = 0                                                # :600-606
        da[np.isinf(da)] = 0
        da_raw = da.copy()
        da = clip_dA(da, A, B_ar[1], mu_ar[1], q_ar[1], threshold=1.0)      # :611
        da = np.clip(da, -1.0, 1.0).reshape(A.shape)
        da = ndimage.median_filter(A + da, size=7) - A                      # :636
        if return_trace:
            trace.append(dict(A=A.copy(), D=D_fwd.copy(), b=b_comb.copy(), A_comb=A_comb.copy(), da_raw=da_raw, da=da.copy(),
                              psi1=psi1, psi2=psi2, psi_c=psi_c, wx=weights_x.copy(), wy=weights_y.copy(),
                              E=(E_f, E_b, E1, E2, E_cons)))
        A = A + da
    return (A, energies, trace) if return_trace else (A, energies)
