"""Host-compiled check of the kernel's ray-march function (racing_dreamer_b200/csrc/rd_march.cuh): the clearance-jump
march must find the SAME hit cell through the SAME side as a plain cell-by-cell DDA, for every ray, on the real maps.
This compiles the header with g++ (tests/native/march_check.cpp); it is a unit test of host logic + integer arithmetic,
not a CPU path of the product."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from racing_dreamer_b200 import load_track

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("march") / "libmarch_check.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                    str(ROOT / "tests" / "native" / "march_check.cpp")], check=True)
    lib = C.CDLL(str(out))
    lib.march_check.restype = C.c_longlong
    lib.march_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong,
                                C.c_void_p]
    lib.clearance_field.restype = None
    lib.clearance_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _rays(tm, n_pose, rng, nb=1080, fov=270.0 * np.pi / 180.0):
    """(px, py, DX, DY) exactly as k_lidar forms them: origin floor(u*4096), direction rint(d*2^18)."""
    p = tm.reset_poses[rng.randint(0, len(tm.reset_poses), n_pose)].copy()
    p[:, :2] += rng.uniform(-0.15, 0.15, (n_pose, 2))
    p[:, 2] = rng.uniform(-np.pi, np.pi, n_pose)
    p[:4, 2] = (0.0, np.pi / 2, np.pi, -np.pi / 2)               # axis-aligned beams (adx == 0 / ady == 0 paths)
    inv = 1.0 / tm.resolution
    PX = np.floor((p[:, 0] - tm.origin[0]) * inv * 4096).astype(np.int64) - tm.c0 * 4096
    PY = np.floor((p[:, 1] - tm.origin[1]) * inv * 4096).astype(np.int64) - tm.cy0 * 4096
    PX[4:40] &= ~4095                                            # origins exactly on a cell edge (bx == 0 / 4096)
    PY[20:60] &= ~4095                                           # ... in y, and on a cell corner for poses 20..39
    bits = np.ascontiguousarray(tm.packed_bits_yup())
    ix, iy = PX >> 12, PY >> 12
    ok = (ix >= 0) & (ix < tm.w) & (iy >= 0) & (iy < tm.h)
    ixc, iyc = np.clip(ix, 0, tm.w - 1), np.clip(iy, 0, tm.h - 1)
    ok &= ((bits[iyc, ixc >> 5] >> (ixc & 31)) & 1).astype(bool)
    p, PX, PY = p[ok], PX[ok], PY[ok]
    ang = 0.5 * fov - np.arange(nb) * (fov / (nb - 1))
    ca, sa = np.cos(ang), np.sin(ang)
    c, s = np.cos(p[:, 2])[:, None], np.sin(p[:, 2])[:, None]
    r = np.empty((len(p), nb, 4), np.int32)
    r[..., 0], r[..., 1] = PX[:, None], PY[:, None]
    r[..., 2] = np.rint((c * ca - s * sa) * 262144)
    r[..., 3] = np.rint((s * ca + c * sa) * 262144)
    return bits, np.ascontiguousarray(r.reshape(-1, 4))


@pytest.mark.parametrize("track", ["austria", "columbia", "treitlstrasse_v2", "barcelona", "gbr", "circle_cw"])
@pytest.mark.parametrize("cshift", [1, 2, 3])
def test_clearance_march_equals_plain_dda(lib, track, cshift):
    tm = load_track(track)
    bits, rays = _rays(tm, 150, np.random.RandomState(7))
    for range_m in (15.0, 2.0):                                   # 2 m: the range limit cuts most rays short
        rsub = int(np.rint(range_m * (1.0 / tm.resolution) * 4096))
        st = (C.c_longlong * 4)()
        bad = lib.march_check(bits.ctypes.data, tm.h, tm.w, tm.row_words(), cshift, rays.ctypes.data, len(rays), rsub, st)
        assert bad == 0, f"{bad} of {len(rays)} rays differ from the cell-by-cell DDA"
    assert st[0] > 0                                              # jumps were actually taken


def test_clearance_field_is_the_chessboard_distance(lib):
    from scipy import ndimage
    tm = load_track("austria")
    bits = np.ascontiguousarray(tm.packed_bits_yup())
    out = np.zeros(tm.h * tm.w, np.uint8)
    ch, cw = C.c_int(), C.c_int()
    lib.clearance_field(bits.ctypes.data, tm.h, tm.w, tm.row_words(), 0, out.ctypes.data, C.byref(ch), C.byref(cw))
    assert (ch.value, cw.value) == (tm.h, tm.w)
    cols = np.arange(tm.w)
    free = ((bits[:, cols >> 5] >> (cols & 31)) & 1).astype(bool)
    want = ndimage.distance_transform_cdt(np.pad(free, 1), metric="chessboard")[1:-1, 1:-1]
    assert np.array_equal(out.reshape(tm.h, tm.w), np.minimum(want, 255))
