"""Seeded synthetic triplets of the shapes BASELINE.json names.

Pure numpy, no GPU, no reference code: this only *generates inputs* (images, flows, rigidity,
homographies, epipoles) with a planted structure field so that the argmin label is known.  The
flows are produced with plain fp64 formulas of the plane+parallax model;
they are inputs, so they need not reproduce OpenCV's rounding.
"""
import numpy as np

SHAPES = {
    "sintel": (436, 1024),     # BASELINE.json configs[1]
    "kitti": (375, 1242),      # configs[2]
    "4k": (2160, 3840),        # configs[3]
}


class Texture:
    """T(x,y) = 0.5 + sum_k a_k sin(2 pi (fx_k x + fy_k y) + phi_k) per RGB channel; evaluable at
    continuous coordinates so that warped frames satisfy brightness constancy exactly."""

    def __init__(self, rs, n_waves=32):
        lam = np.exp(rs.uniform(np.log(6.0), np.log(200.0), size=(3, n_waves)))
        ang = rs.uniform(0, 2 * np.pi, size=(3, n_waves))
        self.fx = np.cos(ang) / lam
        self.fy = np.sin(ang) / lam
        self.phi = rs.uniform(0, 2 * np.pi, size=(3, n_waves))
        self.amp = rs.uniform(0.2, 1.0, size=(3, n_waves)) * (1.6 / n_waves)

    def __call__(self, x, y):
        out = np.empty(x.shape + (3,), np.float64)
        for c in range(3):
            acc = np.full(x.shape, 0.5)
            for k in range(self.fx.shape[1]):
                acc += self.amp[c, k] * np.sin(2 * np.pi * (self.fx[c, k] * x + self.fy[c, k] * y) + self.phi[c, k])
            out[..., c] = acc
        return np.clip(out, 0.0, 1.0)


def _warp_points(A, q, mu, H, b, h, w, x=None, y=None):
    """Where each reference pixel (or continuous point x, y) lands in the neighbour frame for
    structure A (fp64 model; ``A`` an array or a callable A(x, y))."""
    if x is None:
        y, x = np.mgrid[:h, :w].astype(np.float64)
    if callable(A):
        A = A(x, y)
    d = np.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2)
    n = d / mu
    r = (b * A * n) / (b * A / mu - 1.0)
    dd = np.maximum(1e-3, d)
    xn = x + r * (q[0] - x) / dd
    yn = y + r * (q[1] - y) / dd
    den = xn * H[2, 0] + yn * H[2, 1] + H[2, 2]
    X = (xn * H[0, 0] + yn * H[0, 1] + H[0, 2]) / den
    Y = (xn * H[1, 0] + yn * H[1, 1] + H[1, 2]) / den
    return X, Y, x, y


def _inverse_warp(A_fn, q, mu, H, b, h, w, iters=6):
    """Points p of the reference frame with w(p) = p' for every pixel p' of the neighbour grid
    (fixed-point iteration p <- p' - (w(p) - p); the displacement field is smooth, so it contracts)."""
    yp, xp = np.mgrid[:h, :w].astype(np.float64)
    x, y = xp.copy(), yp.copy()
    for _ in range(iters):
        X, Y, _, _ = _warp_points(A_fn, q, mu, H, b, h, w, x=x, y=y)
        x, y = x + (xp - X), y + (yp - Y)
    return x, y


def make_triplet(shape="sintel", seed=1001, epipole="outside", moving_objects=False, flow_noise=0.0,
                 rigidity_value=0.9, n_waves=32, B=(-4.0, 4.0)):
    """Return a dict with the inputs of compute_structure / build_cost_volume / combine_flow.

    ``shape`` is a key of SHAPES or an (h, w) tuple.  ``epipole`` 'outside' puts the FoE at
    (1.4 w, 0.35 h) so correct_epipole returns early; 'inside' at (0.52 w, 0.45 h) to exercise
    the 1e-3 / 0.5 clamps.  ``moving_objects`` pastes independently moving rectangles/ellipses
    (rigidity 0.1 inside) into both flows.
    """
    h, w = SHAPES[shape] if isinstance(shape, str) else shape
    rs = np.random.RandomState(seed)
    tex = Texture(rs, n_waves)

    def homography():
        a, b, c, d = rs.uniform(-2e-3, 2e-3, 4)
        tx, ty = rs.uniform(-3, 3, 2)
        g, hh = rs.uniform(-2e-6, 2e-6, 2) * (1024.0 / w)
        return np.array([[1 + a, b, tx], [c, 1 + d, ty], [g, hh, 1.0]])

    H_fwd = homography()
    H_bwd = np.linalg.inv(H_fwd) @ (np.eye(3) + np.array([[0, 0, rs.uniform(-.5, .5)],
                                                          [0, 0, rs.uniform(-.5, .5)], [0, 0, 0]]))
    H_bwd /= H_bwd[2, 2]
    if epipole == "outside":
        q1 = np.array([1.4 * w, 0.35 * h])
    else:
        q1 = np.array([0.52 * w, 0.45 * h])
    q0 = q1 + rs.uniform(-2, 2, 2)
    y, x = np.mgrid[:h, :w].astype(np.float64)
    mu = [np.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2).mean() for q in (q0, q1)]
    B = [float(B[0]), float(B[1])]

    def A_fn(xx, yy):
        return 1.0 + 0.6 * np.sin(xx / 130.0) * np.cos(yy / 95.0) + 0.002 * (yy - h / 2.0)

    def A_bwd_fn(xx, yy):
        return A_fn(xx, yy) * mu[0] / mu[1]

    A_true = A_fn(x, y)
    A_bwd = A_true * mu[0] / mu[1]

    X1, Y1, _, _ = _warp_points(A_true, q1, mu[1], H_fwd, B[1], h, w)
    X0, Y0, _, _ = _warp_points(A_bwd, q0, mu[0], H_bwd, B[0], h, w)
    flow_f = [X1 - x, Y1 - y]
    flow_b = [X0 - x, Y0 - y]

    rigidity = np.full((h, w), float(rigidity_value))
    if moving_objects:
        for _ in range(6):
            cy, cx = rs.uniform(0.1, 0.9) * h, rs.uniform(0.1, 0.9) * w
            ry, rx = rs.uniform(0.05, 0.12) * h, rs.uniform(0.05, 0.12) * w
            if rs.rand() < 0.5:
                m = ((x - cx) / rx) ** 2 + ((y - cy) / ry) ** 2 < 1
            else:
                m = (np.abs(x - cx) < rx) & (np.abs(y - cy) < ry)
            rigidity[m] = 0.1
            du, dv = rs.uniform(-8, 8, 2)
            for fl, sgn in ((flow_f, 1.0), (flow_b, -1.0)):
                fl[0] = np.where(m, sgn * du, fl[0])
                fl[1] = np.where(m, sgn * dv, fl[1])
    if flow_noise > 0:
        for fl in (flow_f, flow_b):
            fl[0] = fl[0] + rs.normal(0, flow_noise, (h, w))
            fl[1] = fl[1] + rs.normal(0, flow_noise, (h, w))

    # Frames: I1 is the texture itself; a neighbour frame shows at p' the texture of the reference
    # point p that the (rigid) model maps there, I_f(w_f(p)) = T(p): exact brightness constancy for
    # the planted structure, so the data cost has its minimum at the planted label.
    I1 = tex(x, y)
    xi, yi = _inverse_warp(A_fn, q1, mu[1], H_fwd, B[1], h, w)
    I2 = tex(xi, yi)
    xi, yi = _inverse_warp(A_bwd_fn, q0, mu[0], H_bwd, B[0], h, w)
    I0 = tex(xi, yi)

    occ = [np.zeros((h, w), bool), np.zeros((h, w), bool)]
    return dict(
        images=[I0, I1, I2],
        flow=[[flow_b[0].astype(np.float32), flow_b[1].astype(np.float32)],
              [flow_f[0].astype(np.float32), flow_f[1].astype(np.float32)]],
        rigidity=rigidity, occlusions=occ,
        homographies=[H_bwd, H_fwd], epipoles=[q0, q1],
        A_true=A_true, mu=mu, B=B, shape=(h, w), seed=seed,
    )


def make_triplet_device(shape="4k", seed=1001, epipole="outside", rigidity_value=0.9, n_waves=32, B=(-4.0, 4.0),
                        device="cuda"):
    """``make_triplet`` (no moving objects, no flow noise) evaluated with torch on ``device``: the same
    scene parameters drawn from the same numpy stream, the per-pixel fp64 formulas on the GPU, so a 4K
    triplet takes a fraction of a second instead of a minute and a half of host time.  Returns the same
    dict with torch tensors (images fp64 [h,w,3], flows fp32, rigidity fp64) plus host scalars.  The
    pixel values agree with the numpy generator up to the last bits of sin(); they are *inputs*."""
    import torch
    h, w = SHAPES[shape] if isinstance(shape, str) else shape
    rs = np.random.RandomState(seed)
    tex = Texture(rs, n_waves)

    def homography():
        a, b, c, d = rs.uniform(-2e-3, 2e-3, 4)
        tx, ty = rs.uniform(-3, 3, 2)
        g, hh = rs.uniform(-2e-6, 2e-6, 2) * (1024.0 / w)
        return np.array([[1 + a, b, tx], [c, 1 + d, ty], [g, hh, 1.0]])

    H_fwd = homography()
    H_bwd = np.linalg.inv(H_fwd) @ (np.eye(3) + np.array([[0, 0, rs.uniform(-.5, .5)],
                                                          [0, 0, rs.uniform(-.5, .5)], [0, 0, 0]]))
    H_bwd /= H_bwd[2, 2]
    q1 = np.array([1.4 * w, 0.35 * h]) if epipole == "outside" else np.array([0.52 * w, 0.45 * h])
    q0 = q1 + rs.uniform(-2, 2, 2)
    F64 = torch.float64
    y, x = torch.meshgrid(torch.arange(h, dtype=F64, device=device), torch.arange(w, dtype=F64, device=device), indexing="ij")
    mu = [float(torch.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2).mean()) for q in (q0, q1)]
    B = [float(B[0]), float(B[1])]

    def A_fn(xx, yy):
        return 1.0 + 0.6 * torch.sin(xx / 130.0) * torch.cos(yy / 95.0) + 0.002 * (yy - h / 2.0)

    def warp(A, q, m, H, b, xx, yy):
        d = torch.sqrt((xx - q[0]) ** 2 + (yy - q[1]) ** 2)
        r = (b * A * (d / m)) / (b * A / m - 1.0)
        dd = torch.clamp(d, min=1e-3)
        xn = xx + r * (q[0] - xx) / dd
        yn = yy + r * (q[1] - yy) / dd
        den = xn * H[2, 0] + yn * H[2, 1] + H[2, 2]
        return (xn * H[0, 0] + yn * H[0, 1] + H[0, 2]) / den, (xn * H[1, 0] + yn * H[1, 1] + H[1, 2]) / den

    def texture(xx, yy):
        out = torch.empty(xx.shape + (3,), dtype=F64, device=device)
        for c in range(3):
            acc = torch.full_like(xx, 0.5)
            for k in range(tex.fx.shape[1]):
                acc += float(tex.amp[c, k]) * torch.sin(2 * np.pi * (float(tex.fx[c, k]) * xx + float(tex.fy[c, k]) * yy)
                                                        + float(tex.phi[c, k]))
            out[..., c] = acc
        return out.clamp_(0.0, 1.0)

    def inverse_warp(scale, q, m, H, b, iters=6):
        xx, yy = x.clone(), y.clone()
        for _ in range(iters):
            X, Y = warp(A_fn(xx, yy) * scale, q, m, H, b, xx, yy)
            xx, yy = xx + (x - X), yy + (y - Y)
        return xx, yy

    A_true = A_fn(x, y)
    s_b = mu[0] / mu[1]
    X1, Y1 = warp(A_true, q1, mu[1], H_fwd, B[1], x, y)
    X0, Y0 = warp(A_true * s_b, q0, mu[0], H_bwd, B[0], x, y)
    I1 = texture(x, y)
    I2 = texture(*inverse_warp(1.0, q1, mu[1], H_fwd, B[1]))
    I0 = texture(*inverse_warp(s_b, q0, mu[0], H_bwd, B[0]))
    F32 = torch.float32
    return dict(images=[I0, I1, I2],
                flow=[[(X0 - x).to(F32), (Y0 - y).to(F32)], [(X1 - x).to(F32), (Y1 - y).to(F32)]],
                rigidity=torch.full((h, w), float(rigidity_value), dtype=F64, device=device),
                homographies=[H_bwd, H_fwd], epipoles=[q0, q1], A_true=A_true, mu=mu, B=B, shape=(h, w), seed=seed)
