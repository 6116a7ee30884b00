"""Pin the numpy restatements of the OpenCV primitives (oracle/cvemul.py) against OpenCV itself."""
import numpy as np
import cv2
import pytest

from oracle import cvemul


@pytest.mark.parametrize("mode,flag", [("cubic", cv2.INTER_CUBIC), ("linear", cv2.INTER_LINEAR)])
@pytest.mark.parametrize("cn", [1, 3])
def test_remap_bit_exact(mode, flag, cn):
    rs = np.random.RandomState(3)
    h, w = 97, 131
    I = (rs.rand(h, w, cn) * 255).astype(np.float32)
    if cn == 1:
        I = I[..., 0]
    x = (rs.rand(h, w) * (w + 20) - 10).astype(np.float32)
    y = (rs.rand(h, w) * (h + 20) - 10).astype(np.float32)
    # exact grid points, half-way quantisation points and far-out / non-finite coordinates
    x[0, :8] = [0, 1, w - 1, w - 2, 5.015625, 5.046875, -0.015625, w - 1 + 0.015625]
    y[0, :8] = [0, h - 1, 0, h - 2, 7.015625, 7.046875, -0.015625, h - 1]
    x[1, :6] = [1e9, -1e9, np.nan, np.inf, -np.inf, 70000.0]
    y[1, :6] = [3.0, 4.0, 5.0, 6.0, 7.0, -70000.0]
    ref = cv2.remap(I, x, y, interpolation=flag, borderMode=cv2.BORDER_REPLICATE)
    emu = cvemul.remap(I, x, y, mode)
    assert np.array_equal(ref, emu)


def test_perspective_transform_bit_exact():
    rs = np.random.RandomState(4)
    pts = (rs.rand(200000, 2) * [3840, 2160]).astype(np.float32)
    M = np.array([[1.001, 2e-3, 3.1], [-1.5e-3, 0.998, -2.2], [1.5e-6, -1e-6, 1.0]])
    ref = cv2.perspectiveTransform(pts[:, None, :], M).reshape(-1, 2)
    assert np.array_equal(ref, cvemul.perspective_transform(pts, M))
    Mi = np.linalg.inv(M)
    ref = cv2.perspectiveTransform(pts[:, None, :], Mi).reshape(-1, 2)
    assert np.array_equal(ref, cvemul.perspective_transform(pts, Mi))


def test_gray_and_blur_bit_exact():
    rs = np.random.RandomState(5)
    I = rs.rand(61, 83, 3)
    g = cv2.cvtColor(I.astype("float32"), cv2.COLOR_RGB2GRAY)
    assert np.array_equal(g, cvemul.rgb2gray32(I.astype(np.float32)))
    a = rs.standard_normal((61, 83)) * 40
    assert np.array_equal(cv2.GaussianBlur(a, ksize=(3, 3), sigmaX=-1), cvemul.gaussian_blur3(a))


def test_identity_sampling_known_answer():
    # cubic weights at fraction 0 are (0,1,0,0) -> integer grid returns I.
    rs = np.random.RandomState(6)
    I = (rs.rand(20, 30) * 255).astype(np.float32)
    y, x = np.mgrid[:20, :30].astype(np.float32)
    assert np.array_equal(cvemul.remap(I, x, y, "cubic"), I)
    assert np.array_equal(cvemul.remap(I, x, y, "linear"), I)


@pytest.mark.parametrize("cn", [1, 2])
def test_remap_linear_constant_border_bit_exact(cn):
    """cv2.remap's default border (BORDER_CONSTANT, 0) with INTER_LINEAR, as rigidity/occlusions.py uses it."""
    rs = np.random.RandomState(13)
    h, w = 53, 71
    I = (rs.standard_normal((h, w, cn)) * 7).astype(np.float32)
    if cn == 1:
        I = I[..., 0]
    x = (rs.rand(h, w) * (w + 12) - 6).astype(np.float32)
    y = (rs.rand(h, w) * (h + 12) - 6).astype(np.float32)
    x[0, :8] = [0, -0.5, -1, -1.015625, w - 1, w - 0.5, w, w - 1 + 0.015625]
    y[0, :8] = [0, 3, 4, 5, h - 1, h - 0.5, h, -1.0]
    x[1, :5] = [1e9, -1e9, np.nan, np.inf, 70000.0]
    y[1, :5] = [3.0, 4.0, 5.0, 6.0, -70000.0]
    ref = cv2.remap(I, x, y, interpolation=cv2.INTER_LINEAR)
    emu = cvemul.remap(I, x, y, "linear", border="constant")
    assert np.array_equal(ref, emu)


@pytest.mark.parametrize("shape", [(61, 83, 3), (120, 200, 1)])
def test_warp_perspective_and_ptransform64_bit_exact(shape):
    """cv2.warpPerspective(fp64, INTER_LINEAR, BORDER_REPLICATE) and cv2.perspectiveTransform(fp64), the
    primitives of compute_derivs_through_H (structure_refinement/interpolate_through_H.py:53-97)."""
    rs = np.random.RandomState(17)
    h, w, cn = shape
    I = rs.rand(h, w, cn) * 255
    if cn == 1:
        I = I[..., 0]
    H = np.array([[1.002, -0.001, 2.3], [0.0015, 0.999, -1.7], [1.5e-6, -1e-6, 1.0]])
    Hinv = np.linalg.inv(H)
    y, x = np.mgrid[:h, :w]
    pts = np.c_[x.ravel(), y.ravel()].astype("float")
    for T in (np.array([[1.0, 0, -1.0], [0, 1, 0], [0, 0, 1]]), np.array([[1.0, 0, 0], [0, 1, 1.0], [0, 0, 1]]), np.eye(3)):
        for M in (Hinv.dot(T).dot(H), T):
            ref = cv2.warpPerspective(I, M, (w, h), borderMode=cv2.BORDER_REPLICATE)
            assert np.array_equal(ref, cvemul.warp_perspective(I, M))
            sub = pts[::37]
            refp = cv2.perspectiveTransform(sub[:, None, :], M).squeeze()
            assert np.array_equal(refp, cvemul.perspective_transform64(sub, M))
