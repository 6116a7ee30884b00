"""Plain-numpy, bit-exact restatements of the OpenCV primitives on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

OpenCV is a third-party dependency of the reference (not vendored under /root/reference; the
reference pins no version, README.md:19).  Call sites on the path:
``utils/flow_homography.py:15,36,46`` (perspectiveTransform),
``structure_refinement/interpolation.py:76-95`` (remap, INTER_CUBIC, BORDER_REPLICATE),
``structure_refinement/optimize_structure_single_scale.py:352,362-369`` (cvtColor, GaussianBlur),
``rigidity/occlusions.py`` (remap, INTER_LINEAR).  The semantics below were pinned empirically
against OpenCV 4.13.0 in this image (``tests/test_oracle_cvemul.py`` asserts bit equality) and
are the arithmetic specification the CUDA kernels implement.
"""
import numpy as np

F32 = np.float32
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INT_MIN = -(1 << 31)


def cubic_table():
    """32x4 fp32 bicubic weights, A=-0.75 (OpenCV ``interpolateCubic`` evaluated at i/32)."""
    A = F32(-0.75)
    one = F32(1)
    t = np.arange(INTER_TAB_SIZE, dtype=F32) * F32(1.0 / INTER_TAB_SIZE)
    c0 = ((A * (t + one) - F32(5) * A) * (t + one) + F32(8) * A) * (t + one) - F32(4) * A
    c1 = ((A + F32(2)) * t - (A + F32(3))) * t * t + one
    u = one - t
    c2 = ((A + F32(2)) * u - (A + F32(3))) * u * u + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], axis=1).astype(F32)


def linear_table():
    """32x2 fp32 bilinear weights (OpenCV ``interpolateLinear``)."""
    t = np.arange(INTER_TAB_SIZE, dtype=F32) * F32(1.0 / INTER_TAB_SIZE)
    return np.stack([F32(1) - t, t], axis=1).astype(F32)


def quantise(coord32):
    """``cvRound(coord*32)`` as cv2.remap does it: fp32 product, round-half-even, and the x86
    "integer indefinite" value for NaN / out-of-int32-range inputs.  Returns (int part, frac)."""
    with np.errstate(invalid="ignore", over="ignore"):
        v = (coord32.astype(F32) * F32(INTER_TAB_SIZE)).astype(F32)
        ok = np.abs(v) < F32(2147483648.0)
        s = np.where(ok, np.rint(np.where(ok, v, 0)), INT_MIN).astype(np.int64)
    ipart = np.clip(s >> INTER_BITS, -32768, 32767)  # saturate_cast<short>
    return ipart, s & (INTER_TAB_SIZE - 1)


def remap(img32, x32, y32, mode="cubic", border="replicate"):
    """cv2.remap(img fp32, mapx fp32, mapy fp32, INTER_CUBIC|INTER_LINEAR, BORDER_REPLICATE), or with
    ``border='constant'`` (linear only) BORDER_CONSTANT with value 0, cv2.remap's default, which is what
    rigidity/occlusions.py:33-40 uses: taps outside the image contribute 0.

    fp32 throughout; 2-D weight = fp32(wy*wx).  Summation order is OpenCV's.  Cubic: for a
    footprint fully inside the image each row is summed left to right and the four row sums are
    then accumulated; for a footprint touching the border all taps are accumulated one by one.
    Linear: always tap by tap.
    """
    img32 = np.asarray(img32, dtype=F32)
    squeeze = img32.ndim == 2
    if squeeze:
        img32 = img32[:, :, None]
    h, w = img32.shape[:2]
    ix, fx = quantise(np.asarray(x32, F32))
    iy, fy = quantise(np.asarray(y32, F32))
    if mode == "cubic":
        tab, k, off = cubic_table(), 4, 1
    elif mode == "linear":
        tab, k, off = linear_table(), 2, 0
    else:
        raise ValueError(mode)
    wx, wy = tab[fx], tab[fy]
    ix0, iy0 = ix - off, iy - off
    inside = (ix0 >= 0) & (ix0 < max(w - (k - 1), 0)) & (iy0 >= 0) & (iy0 < max(h - (k - 1), 0))
    seq = None   # border path: tap-by-tap
    grp = None   # interior path: row sums, then rows
    if border == "constant" and mode != "linear":
        raise ValueError("constant border is restated for the linear mode only")
    for j in range(k):
        yy = np.clip(iy0 + j, 0, h - 1)
        row = None
        for i in range(k):
            xx = np.clip(ix0 + i, 0, w - 1)
            wgt = (wy[..., j] * wx[..., i]).astype(F32)
            tap = img32[yy, xx]
            if border == "constant":
                out_of_image = ((iy0 + j) < 0) | ((iy0 + j) >= h) | ((ix0 + i) < 0) | ((ix0 + i) >= w)
                tap = np.where(out_of_image[..., None], F32(0), tap)
            p = (tap * wgt[..., None]).astype(F32)
            row = p if row is None else (row + p).astype(F32)
            seq = p if seq is None else (seq + p).astype(F32)
        grp = row if grp is None else (grp + row).astype(F32)
    out = (np.where(inside[..., None], grp, seq) if mode == "cubic" else seq).astype(F32)
    return out[..., 0] if squeeze else out


def perspective_transform(pts32, M):
    """cv2.perspectiveTransform on an [N,2] fp32 array with a 3x3 fp64 matrix: fp64 arithmetic
    (no FMA), ``w = 1/w`` then multiply, one rounding to fp32; ``|w| <= FLT_EPSILON`` -> 0."""
    M = np.asarray(M, dtype=np.float64)
    x = pts32[:, 0].astype(np.float64)
    y = pts32[:, 1].astype(np.float64)
    wv = (x * M[2, 0] + y * M[2, 1]) + M[2, 2]
    ok = np.abs(wv) > np.finfo(F32).eps
    with np.errstate(divide="ignore", invalid="ignore"):
        iw = np.where(ok, 1.0 / np.where(ok, wv, 1.0), 0.0)
    X = ((x * M[0, 0] + y * M[0, 1]) + M[0, 2]) * iw
    Y = ((x * M[1, 0] + y * M[1, 1]) + M[1, 2]) * iw
    return np.stack([X, Y], axis=1).astype(F32)


def _fma32(a, b, c):
    # fp32 fused multiply-add emulated in fp64 (exact product of two fp32 fits in fp64; the single
    # rounding of the sum to fp64 then fp32 can double-round only in cases that did not occur on
    # any tested input; the test asserts equality with cv2).
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(F32)


def rgb2gray32(rgb32):
    """cv2.cvtColor(fp32 RGB, COLOR_RGB2GRAY) of OpenCV 4.13 on this image's CPUs:
    ``fma(B, 0.114f, fma(R, 0.299f, G*0.587f))``."""
    r, g, b = rgb32[..., 0], rgb32[..., 1], rgb32[..., 2]
    return _fma32(b, F32(0.114), _fma32(r, F32(0.299), (g * F32(0.587)).astype(F32)))


def gaussian_blur3(a):
    """cv2.GaussianBlur(fp64, ksize=(3,3), sigmaX=-1): separable [1/4,1/2,1/4], BORDER_REFLECT_101.
    Row pass sums left to right, column pass adds the two outer rows first (pinned empirically)."""
    p = np.pad(np.asarray(a, np.float64), 1, mode="reflect")
    r = (p[:, :-2] * 0.25 + p[:, 1:-1] * 0.5) + p[:, 2:] * 0.25
    return r[1:-1] * 0.5 + (r[:-2] + r[2:]) * 0.25


def invert3x3(M):
    """cv::invert of a 3x3 CV_64F matrix (closed form: cofactors times 1/det, OpenCV's operation order)."""
    S = np.asarray(M, dtype=np.float64)
    det = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0])
           + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    d = 1.0 / det
    t = np.empty(9)
    t[0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    t[1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    t[2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    t[3] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    t[4] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    t[5] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    t[6] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    t[7] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    t[8] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t.reshape(3, 3)


def warp_perspective(img64, M, border="replicate"):
    """cv2.warpPerspective(fp64 image, M, (w, h), INTER_LINEAR, BORDER_REPLICATE) -- the call of
    structure_refinement/interpolate_through_H.py:90-93.  OpenCV inverts M (closed form), walks the
    destination in blocks, maps each pixel as ((X0 + m0*x1) * (32/W)) rounded to an integer with 5
    fraction bits, and interpolates bilinearly with fp32 table weights in fp64."""
    img64 = np.asarray(img64, dtype=np.float64)
    squeeze = img64.ndim == 2
    if squeeze:
        img64 = img64[:, :, None]
    h, w = img64.shape[:2]
    Mi = invert3x3(M).ravel()
    BLOCK = 32
    bh0 = min(BLOCK // 2, h)
    bw0 = min(BLOCK * BLOCK // bh0, w)
    bh0 = min(BLOCK * BLOCK // bw0, h)
    yy, xx = np.mgrid[:h, :w]
    bx = (xx // bw0) * bw0
    x1 = (xx - bx).astype(np.float64)
    bx = bx.astype(np.float64)
    yf = yy.astype(np.float64)
    X0 = Mi[0] * bx + Mi[1] * yf + Mi[2]
    Y0 = Mi[3] * bx + Mi[4] * yf + Mi[5]
    W0 = Mi[6] * bx + Mi[7] * yf + Mi[8]
    with np.errstate(divide="ignore", invalid="ignore"):
        W = W0 + Mi[6] * x1
        W = np.where(W != 0, INTER_TAB_SIZE / W, 0.0)
        fX = np.maximum(float(INT_MIN), np.minimum(2147483647.0, (X0 + Mi[0] * x1) * W))
        fY = np.maximum(float(INT_MIN), np.minimum(2147483647.0, (Y0 + Mi[3] * x1) * W))
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = X & (INTER_TAB_SIZE - 1)
    fy = Y & (INTER_TAB_SIZE - 1)
    tab = linear_table()
    wx, wy = tab[fx], tab[fy]
    out = None
    for j in range(2):
        for i in range(2):
            wgt = (wy[..., j] * wx[..., i]).astype(F32).astype(np.float64)
            yc = np.clip(sy + j, 0, h - 1)
            xc = np.clip(sx + i, 0, w - 1)
            p = img64[yc, xc] * wgt[..., None]
            out = p if out is None else out + p
    return out[..., 0] if squeeze else out


def _fma(a, b, c):
    """exact fused multiply-add of python floats (rational arithmetic, one rounding)"""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def perspective_transform64(pts64, M):
    """cv2.perspectiveTransform on an [N,2] fp64 array (interpolate_through_H.py:73-79 passes fp64
    points).  OpenCV 4.13's fp64 kernel is compiled with FMA contraction: each row is
    ``fma(x, m0, y*m1) + m2``, then ``w = 1/w`` and a multiply (pinned empirically; the fp32 kernel above is
    not contracted).  Scalar loop -- for small N only."""
    m = np.asarray(M, dtype=np.float64).ravel()
    out = np.empty((len(pts64), 2))
    for i, (x, y) in enumerate(np.asarray(pts64, dtype=np.float64)):
        x, y = float(x), float(y)
        w = _fma(x, m[6], y * m[7]) + m[8]
        iw = 1.0 / w if abs(w) > np.finfo(np.float64).eps else 0.0
        out[i, 0] = (_fma(x, m[0], y * m[1]) + m[2]) * iw
        out[i, 1] = (_fma(x, m[3], y * m[4]) + m[5]) * iw
    return out
