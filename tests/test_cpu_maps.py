"""CPU: compiled tracks (a6) -- format, geometry conventions, and the generator restatement."""
import numpy as np
import pytest

from oracle import ref_stubs
from racing_dreamer_b200 import TRACK_FILES, load_track, maps

# lap lengths of the paper's tracks: wavefront distance * 0.05 m
LAP_M = {"austria": 79.45, "barcelona": 201.0, "treitlstrasse_v2": 51.65, "columbia": 61.2, "circle_cw": 41.9}


@pytest.mark.parametrize("name", sorted(TRACK_FILES))
def test_compiled_track(name):
    tm = load_track(name)
    assert tm.resolution == 0.05 and tm.full_shape == (2000, 2000)
    if name in LAP_M:
        assert abs(tm.lap_length_m() - LAP_M[name]) < 0.051
    d = tm.drivable
    assert not d[:maps.MARGIN].any() and not d[-maps.MARGIN:].any() and not d[:, :maps.MARGIN].any() and not d[:, -maps.MARGIN:].any()
    assert tm.dist.max() == tm.dmax and (tm.dist[~d] == 0).all()
    bits = tm.packed_bits_yup()
    assert bits.shape == (tm.h, tm.row_words()) and tm.row_words() % 2 == 1
    un = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(tm.h, -1)[:, :tm.w].astype(bool)
    assert np.array_equal(un, d[::-1])
    # start pixel = world (0,0) -> (row 999, col 1000) [REF docs/maps/costmaps/generate-costmap.py:49-52]
    assert tm.to_pixel(0.0, 0.0) == (999, 1000)
    r, c = tm.to_pixel(*tm.start_poses[0, :2])
    assert d[r - tm.r0, c - tm.c0] and tm.dist[r - tm.r0, c - tm.c0] == 0
    # every reset pose sits on a drivable cell with clearance
    for x, y, _ in tm.reset_poses[:: max(1, len(tm.reset_poses) // 64)]:
        r, c = tm.to_pixel(x, y)
        assert d[r - tm.r0, c - tm.c0] and tm.edt_sq[r - tm.r0, c - tm.c0] >= 64


def test_reference_format_arrays():
    tm = load_track("treitlstrasse_v2")
    full = tm.full_drivable()
    assert full.shape == (2000, 2000) and full.sum() == tm.drivable.sum()
    nd = tm.full_norm_distance_from_start()
    assert nd.max() == 1.0 and nd.min() == 0.0
    no = tm.full_norm_distance_to_obstacle()
    assert no.max() == 1.0 and (no[~full] == 0).all()


@pytest.mark.skipif(not ref_stubs.available(), reason="/root/reference not present (GPU box)")
def test_recompile_matches_stored_and_dilation_wavefront():
    """Recompiling from the reference's map files reproduces the stored npz, and the frontier BFS equals the
    generator's literal repeated-3x3-dilation wavefront [REF docs/maps/costmaps/generate-costmap.py:198-209]."""
    from scipy import ndimage
    tm = load_track("treitlstrasse_v2")
    fresh = maps.compile_track(ref_stubs.REFERENCE_ROOT / "docs/maps/maps/Treitlstrasse_3-U_v2.yaml")
    assert np.array_equal(fresh.drivable, tm.drivable) and np.array_equal(fresh.dist, tm.dist)
    assert np.array_equal(fresh.reset_poses, tm.reset_poses)
    free = tm.drivable & (tm.dist < tm.dmax)
    r, c = tm.to_pixel(0.0, 0.0)
    mask = np.zeros_like(free)
    mask[r - tm.r0, c - tm.c0] = True
    dist = np.zeros(free.shape, np.int32)
    cur = 0
    while True:
        cur += 1
        new = free & (ndimage.binary_dilation(mask, structure=np.ones((3, 3), bool)) ^ mask)
        if not new.any():
            break
        dist[new] = cur
        mask |= new
    assert cur == tm.dmax
    assert np.array_equal(dist[free], tm.dist[free].astype(np.int32))
