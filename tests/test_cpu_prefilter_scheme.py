"""Error bound of k_occupancy's float32 prefilter scheme (csrc/rd_occupancy.cuh occ_prefilter_axis), emulated in NumPy.

The kernel filters a 220-sample line with 32 lanes of 7 samples (4 zero slots in front of sample 0), zero incoming
state per lane, neighbours' end values restoring the true state, inputs scaled by -z.  This test restates that scheme
operation by operation in float32 (fma = one rounding, emulated through float64) and compares it with scipy's float64
spline_filter1d on binary lines and on a second pass over real-valued lines: the kernel's exactness stage assumes the
float32 coefficient image is within 2e-6 of the float64 one (OCC_EPS = 2e-5 leaves a factor 10).  Host logic only."""
import numpy as np
import pytest
from scipy import ndimage

f32 = np.float32
Z = f32(-0.26794919243112270647)


def fma(a, b, c):
    return (np.float64(a) * np.float64(b) + np.float64(c)).astype(np.float32)


def lane_scheme(lines):
    """lines: [m, 220] float32 inputs (not yet multiplied by the gain).  Returns the prefiltered lines (float32)."""
    m, n = lines.shape
    assert n == 220
    z = Z
    zp = [z]
    for _ in range(6):
        zp.append(f32(zp[-1] * z))
    z4, z5, z7 = zp[3], zp[4], zp[6]
    zm5 = f32(1.0) / z5
    zm9 = f32(1.0) / f32(z5 * z4)
    xg = f32(-z * f32(6.0))
    x = np.zeros((m, 32, 7), np.float32)
    padded = np.concatenate([np.zeros((m, 4), np.float32), lines.astype(np.float32)], axis=1)      # 224 = 32 * 7
    x[:] = (padded * xg).astype(np.float32).reshape(m, 32, 7)
    u = np.zeros_like(x)
    u[..., 0] = x[..., 0]
    for j in range(1, 7):
        u[..., j] = fma(z, u[..., j - 1], x[..., j])
    poly = x[..., 6].copy()
    for j in range(5, -1, -1):
        poly = fma(z, poly, x[..., j])
    m4 = fma(z7, fma(z7, fma(z7, poly[:, 3], poly[:, 2]), poly[:, 1]), poly[:, 0])
    c = fma(m4, zm9, (x[:, 0, 4] * (-zm5)).astype(np.float32))
    c1 = np.zeros((m, 32), np.float32); c1[:, 1:] = u[:, :-1, 6]; c1[:, 0] = c
    c2 = np.zeros_like(c1); c2[:, 1:] = c1[:, :-1]
    c3 = np.zeros_like(c1); c3[:, 1:] = c2[:, :-1]
    sin = fma(z7, fma(z7, c3, c2), c1)
    y = np.zeros_like(x)
    for j in range(7):
        y[..., j] = fma(zp[j], sin, u[..., j])
    w_end = (fma(z, y[:, 31, 5], y[:, 31, 6]) * f32(-1.0 / (float(z) * float(z) - 1.0))).astype(np.float32)
    v = np.zeros_like(x)
    v[..., 6] = y[..., 6]
    v[:, 31, 6] = w_end
    for j in range(5, -1, -1):
        v[..., j] = fma(z, v[..., j + 1], y[..., j])
    d1 = np.zeros((m, 32), np.float32); d1[:, :-1] = v[:, 1:, 0]
    d2 = np.zeros_like(d1); d2[:, :-1] = d1[:, 1:]
    d3 = np.zeros_like(d1); d3[:, :-1] = d2[:, 1:]
    tin = fma(z7, fma(z7, d3, d2), d1)
    out = np.zeros_like(x)
    for j in range(7):
        out[..., j] = fma(zp[6 - j], tin, v[..., j])
    return out.reshape(m, 224)[:, 4:]


def test_lane_scheme_matches_scipy_on_binary_and_second_pass():
    rng = np.random.RandomState(5)
    crops = []
    for _ in range(12):                                    # blocky binary crops (runs of 3..40 equal cells) + noise crops
        img = np.zeros((220, 220), np.float32)
        for r in range(220):
            c = 0
            val = rng.randint(2)
            while c < 220:
                k = rng.randint(3, 41)
                img[r, c:c + k] = val
                val ^= 1
                c += k
        crops.append(img)
    crops.append((rng.rand(220, 220) < 0.5).astype(np.float32))
    crops.append(np.ones((220, 220), np.float32))
    worst1 = worst2 = 0.0
    for img in crops:
        want1 = ndimage.spline_filter1d(img.astype(np.float64), order=3, axis=1, mode="mirror")
        got1 = lane_scheme(img)
        worst1 = max(worst1, np.abs(got1 - want1).max())
        want2 = ndimage.spline_filter1d(want1, order=3, axis=0, mode="mirror")
        got2 = lane_scheme(np.ascontiguousarray(got1.T)).T
        worst2 = max(worst2, np.abs(got2 - want2).max())
    assert worst1 < 6e-7, worst1                            # one axis
    assert worst2 < 2e-6, worst2                            # both axes: the bound the exactness stage relies on
