"""Track-map compiler: ROS map_server yaml+image -> the grids the env step consumes.

Restates the semantics of the reference's offline generator
[REF docs/maps/costmaps/generate-costmap.py:30-52 (load/threshold/start pixel),
 :131-224 (finish-line blocking + 8-neighbour wavefront), :380-382 (EDT), :405-420 (npz keys)]
and the track-name table of SURVEY.md §8-a6.  It is host-side, offline work (the reference
runs it once per track too); its output is cached as a small ``.npz`` under
``racing_dreamer_b200/data/tracks`` so nothing at run time needs the source images.

Compiled layout (all arrays cropped to the track bounding box + MARGIN cells, image row order):

* ``drivable``  bool (h, w)      = reference ``drivable_area`` (wavefront-reachable + finish line)
* ``dist``      uint16 (h, w)    = wavefront distance in cells (finish-line cells = dmax)
                                   -> ``norm_distance_from_start = dist / dmax``
* ``edt_sq``    uint32 (h, w)    = squared Euclidean distance (cells^2) to the nearest
                                   non-drivable cell -> ``norm_distance_to_obstacle``
* ``start_poses`` / ``reset_poses`` float64 (n, 3) = (x, y, yaw) for reset modes
  'grid' / 'random' [NEW-SPEC: the reference's sampler lives in racecar_gym, not in tree].
"""
from __future__ import annotations

import dataclasses
import math
import os
from pathlib import Path
from typing import Dict, Optional

import numpy as np

MARGIN = 4  # cells of guaranteed non-drivable border around the crop (the ray march relies on >= 1)

DATA_DIR = Path(__file__).resolve().parent / "data" / "tracks"

# racecar_gym track name -> map file stem in docs/maps/maps (lap lengths cross-checked there)
TRACK_FILES: Dict[str, str] = {
    "austria": "f1_aut",
    "barcelona": "f1_esp",
    "gbr": "f1_gbr",
    "treitlstrasse_v2": "Treitlstrasse_3-U_v2",
    "columbia": "columbia_small",
    "circle_cw": "circle",
    # further TU-Wien redraws in docs/maps/maps [REF docs/maps/README.md:22-46] (not racecar_gym scene names)
    "monaco": "f1_mco",
    "austria_wide": "f1_aut_wide",
    "treitlstrasse_v1": "Treitlstrasse_3-U_v1",
    "treitlstrasse_v3": "Treitlstrasse_3-U_v3",
}


@dataclasses.dataclass
class TrackMap:
    name: str
    source: str
    resolution: float
    origin: tuple           # (ox, oy) world coordinates of the lower-left corner of the FULL image
    full_shape: tuple       # (H, W) of the full image
    r0: int                 # crop origin, image row/col in the full image
    c0: int
    drivable: np.ndarray    # bool (h, w)
    dist: np.ndarray        # uint16 (h, w)
    dmax: int
    edt_sq: np.ndarray      # uint32 (h, w)
    edt_sq_max: int
    start_poses: np.ndarray  # f64 (n, 3)
    reset_poses: np.ndarray  # f64 (m, 3)

    # ---- geometry helpers (the one convention used by oracle, kernels and the GridMap shim) ----
    @property
    def h(self) -> int:
        return int(self.drivable.shape[0])

    @property
    def w(self) -> int:
        return int(self.drivable.shape[1])

    @property
    def inv_res(self) -> float:
        return 1.0 / self.resolution

    @property
    def cy0(self) -> int:
        """y-up cell index (in the full image) of the crop's bottom row."""
        return self.full_shape[0] - 1 - (self.r0 + self.h - 1)

    def to_pixel(self, x: float, y: float):
        """(row, col) in the FULL image; row = H-1-floor((y-oy)/res) [REF generate-costmap.py:49-52]."""
        col = math.floor((x - self.origin[0]) * self.inv_res)
        row = self.full_shape[0] - 1 - math.floor((y - self.origin[1]) * self.inv_res)
        return row, col

    def lap_length_m(self) -> float:
        return self.dmax * self.resolution

    # ---- full-size reference-format arrays (what the reference's maps.npz holds) ----
    def _paste(self, crop: np.ndarray, dtype) -> np.ndarray:
        full = np.zeros(self.full_shape, dtype=dtype)
        full[self.r0:self.r0 + self.h, self.c0:self.c0 + self.w] = crop
        return full

    def full_drivable(self) -> np.ndarray:
        return self._paste(self.drivable, bool)

    def full_norm_distance_from_start(self) -> np.ndarray:
        d = self._paste(self.dist, np.float64) * float(self.resolution)
        return d / np.amax(d)

    def full_norm_distance_to_obstacle(self) -> np.ndarray:
        d = np.sqrt(self._paste(self.edt_sq, np.float64)) * float(self.resolution)
        return d / np.amax(d)

    # ---- device layouts ----
    def row_words(self) -> int:
        """32-bit words per bit-packed row: >= ceil(w/32), forced odd so that vertically adjacent
        cells fall into different shared-memory banks."""
        rw = (self.w + 31) // 32
        return rw | 1

    def packed_bits_yup(self) -> np.ndarray:
        """uint32 (h, row_words) bit grid, row 0 = lowest y; bit (cx & 31) of word (cx >> 5)."""
        rw = self.row_words()
        yup = self.drivable[::-1]
        padded = np.zeros((self.h, rw * 32), dtype=np.uint8)
        padded[:, : self.w] = yup
        b = padded.reshape(self.h, rw, 32).astype(np.uint32)
        words = (b << np.arange(32, dtype=np.uint32)[None, None, :]).sum(axis=2, dtype=np.uint64)
        return words.astype(np.uint32)

    def dist_yup(self) -> np.ndarray:
        return np.ascontiguousarray(self.dist[::-1])

    # ---- io ----
    def save(self, path: os.PathLike) -> None:
        np.savez_compressed(
            path,
            name=self.name, source=self.source, resolution=self.resolution,
            origin=np.asarray(self.origin, np.float64), full_shape=np.asarray(self.full_shape, np.int64),
            r0=self.r0, c0=self.c0,
            drivable=np.packbits(self.drivable, axis=1), w=self.w,
            dist=self.dist, dmax=self.dmax, edt_sq=self.edt_sq, edt_sq_max=self.edt_sq_max,
            start_poses=self.start_poses, reset_poses=self.reset_poses,
        )

    @staticmethod
    def load(path: os.PathLike) -> "TrackMap":
        z = np.load(path, allow_pickle=False)
        w = int(z["w"])
        drivable = np.unpackbits(z["drivable"], axis=1)[:, :w].astype(bool)
        return TrackMap(
            name=str(z["name"]), source=str(z["source"]), resolution=float(z["resolution"]),
            origin=tuple(float(v) for v in z["origin"]), full_shape=tuple(int(v) for v in z["full_shape"]),
            r0=int(z["r0"]), c0=int(z["c0"]), drivable=drivable, dist=z["dist"], dmax=int(z["dmax"]),
            edt_sq=z["edt_sq"], edt_sq_max=int(z["edt_sq_max"]),
            start_poses=z["start_poses"], reset_poses=z["reset_poses"],
        )


# --------------------------------------------------------------------------------------------
# compiler
# --------------------------------------------------------------------------------------------
def _read_gray(image_path: Path) -> np.ndarray:
    """Grey image as float64, as ``skimage.io.imread(as_gray=True).astype(float)`` yields it
    [REF generate-costmap.py:39]: RGB(A) -> 0.2125 R + 0.7154 G + 0.0721 B on [0,1] (alpha is 255
    everywhere in the in-tree PNGs, so blending is the identity); single-channel files stay as stored."""
    from PIL import Image

    im = Image.open(image_path)
    a = np.asarray(im)
    if a.ndim == 2:
        return a.astype(np.float64)
    rgb = a[..., :3].astype(np.float64) / 255.0
    if a.shape[2] == 4:
        alpha = a[..., 3:4].astype(np.float64) / 255.0
        rgb = rgb * alpha + (1.0 - alpha)
    return rgb @ np.array([0.2125, 0.7154, 0.0721])


def _wavefront(free: np.ndarray, seed_rc, max_dist: Optional[int] = None) -> np.ndarray:
    """8-neighbour (Chebyshev) wavefront distance from the seed over ``free`` cells
    [REF generate-costmap.py:198-209: repeated 3x3 dilation, new pixels get the current distance].
    Returns int32 distances, -1 where unreached."""
    h, w = free.shape
    dist = np.full((h, w), -1, dtype=np.int32)
    dist[seed_rc] = 0
    fr = np.array([seed_rc[0]], dtype=np.int64)
    fc = np.array([seed_rc[1]], dtype=np.int64)
    d = 0
    offs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    while fr.size and (max_dist is None or d < max_dist):
        d += 1
        nr = np.concatenate([fr + a for a, _ in offs])
        nc = np.concatenate([fc + b for _, b in offs])
        ok = (nr >= 0) & (nr < h) & (nc >= 0) & (nc < w)
        nr, nc = nr[ok], nc[ok]
        ok = free[nr, nc] & (dist[nr, nc] < 0)
        nr, nc = nr[ok], nc[ok]
        if nr.size == 0:
            break
        lin = np.unique(nr * w + nc)
        fr, fc = lin // w, lin % w
        dist[fr, fc] = d
    return dist


# The generator clears one hard-coded pixel of EVERY map it processes ("extra for optimized spline at
# Treitlstrasse_3-U_v3") [REF generate-costmap.py:45-46].  On Treitlstrasse_3-U_v2 that pixel lies on the track, where it
# would be a 5 cm phantom wall for the ray caster, so the shipped tracks are compiled without it; pass
# reference_quirks=True to reproduce the generator's arrays bit for bit (tests/test_cpu_maps.py does).
REFERENCE_CLEARED_PIXEL = (987, 1294)


def compile_track(yaml_path: os.PathLike, name: Optional[str] = None, start_xy=(0.0, 0.0),
                  reset_clearance_m: float = 0.4, max_reset_poses: int = 4096,
                  reference_quirks: bool = False) -> TrackMap:
    import yaml
    from scipy import ndimage

    yaml_path = Path(yaml_path)
    with open(yaml_path) as f:
        props = yaml.safe_load(f)
    res = float(props["resolution"])
    ox, oy = float(props["origin"][0]), float(props["origin"][1])
    gray = _read_gray(yaml_path.parent / props["image"])
    H, W = gray.shape
    binary = (gray / np.amax(gray)) > float(props["occupied_thresh"])     # [REF :42-43]
    if reference_quirks and H > REFERENCE_CLEARED_PIXEL[0] and W > REFERENCE_CLEARED_PIXEL[1]:
        binary[REFERENCE_CLEARED_PIXEL] = False

    # start pixel [REF :49-52]; the reference mixes shape[1] with the row axis (square images only) --
    # here the row flip uses the row count.
    gx = int((start_xy[0] - ox) / res)
    gy = int(H - (start_xy[1] - oy) / res - 1)
    if not binary[gy, gx]:
        raise ValueError(f"start pixel ({gy},{gx}) of {yaml_path.name} is not free")

    # finish line: free cells of the column one behind the start, walked down and up [REF :151-163]
    free = binary.copy()
    finish = np.zeros_like(free)
    col = gx - 1
    r = gy
    while free[r, col]:
        free[r, col] = False
        finish[r, col] = True
        r += 1
    r = gy - 1
    while free[r, col]:
        free[r, col] = False
        finish[r, col] = True
        r -= 1

    dist = _wavefront(free, (gy, gx))
    reached = dist >= 0
    dmax = int(dist.max()) + 1                 # loop-exit value of current_distance [REF :198-220]
    dist = np.where(reached, dist, 0)
    dist[finish] = dmax
    drivable = reached | finish                # [REF :223]
    if dmax > 65535:
        raise ValueError("wavefront distance overflows uint16")

    edt = ndimage.distance_transform_edt(drivable)     # [REF :380]
    edt_sq = np.rint(edt * edt).astype(np.uint32)

    rows = np.flatnonzero(drivable.any(axis=1))
    cols = np.flatnonzero(drivable.any(axis=0))
    r0, r1 = max(rows[0] - MARGIN, 0), min(rows[-1] + MARGIN + 1, H)
    c0, c1 = max(cols[0] - MARGIN, 0), min(cols[-1] + MARGIN + 1, W)
    sl = (slice(r0, r1), slice(c0, c1))

    tm = TrackMap(
        name=name or yaml_path.stem, source=yaml_path.stem, resolution=res, origin=(ox, oy), full_shape=(H, W),
        r0=int(r0), c0=int(c0), drivable=drivable[sl].copy(), dist=dist[sl].astype(np.uint16), dmax=dmax,
        edt_sq=edt_sq[sl].copy(), edt_sq_max=int(edt_sq.max()),
        start_poses=np.zeros((0, 3)), reset_poses=np.zeros((0, 3)),
    )
    tm.start_poses = _grid_poses(tm, start_xy)
    tm.reset_poses = _random_reset_poses(tm, reset_clearance_m, max_reset_poses)
    return tm


# The generator's per-map switch for the smoothing terms [REF generate-costmap.py:84-111]: maps with custom settings that
# turn `use_blurred_factor` off keep the plain wavefront distance; every other map adds the two blurred fields.
_NO_BLUR_MAPS = ("f1_aut", "columbia_small")


# per-map settings of the generator [REF docs/maps/costmaps/generate-costmap.py:83-106]:
# erosion of the track for the race line (pixels), B-spline degree, whether the blurred terms enter the smoothed distance,
# and every how many steps of the greedy descent a control point is taken
_GENERATOR_SETTINGS = {
    "Treitlstrasse_3-U_v3": dict(erosion=9, degree=25, blurred=True, sample_every=10),
    "f1_aut": dict(erosion=12, degree=30, blurred=False, sample_every=10),
    "columbia_small": dict(erosion=27, degree=40, blurred=False, sample_every=30),
}
_GENERATOR_DEFAULTS = dict(erosion=12, degree=25, blurred=True, sample_every=15)


def generator_settings(stem: str) -> Dict[str, object]:
    return dict(_GENERATOR_SETTINGS.get(stem, _GENERATOR_DEFAULTS))


def compile_distance_to_target(yaml_path: os.PathLike, start_xy=(0.0, 0.0), reference_quirks: bool = False,
                               use_blurred_factor: Optional[bool] = None, erosion: int = 0) -> Dict[str, np.ndarray]:
    """The fourth layer of the reference's costmap files, ``norm_distance_to`` = normalised SMOOTHED distance to the
    target (the finish line approached in driving direction) [REF docs/maps/costmaps/generate-costmap.py:227-276
    compute_distance_transform_smoothed(forward_direction=False), saved at :405-420], full image size, float64.
    SURVEY.md §8-f4.  No env-path function reads it (the env uses the forward progress map); it completes the map
    compiler's output for tools that load the reference's ``maps.npz`` keys.  Returns ``{'drivable_area',
    'norm_distance_to', 'distance_to'}`` (the latter un-normalised).  ``erosion`` > 0 shrinks the free space by a disk of
    that radius first [REF :135-138]: the variant the race line is traced on (`compile_raceline`).

    Restated steps: backward wavefront (the finish line is blocked one column AFTER the start [REF :160]); 'start' /
    'target' areas = cells within 100 (small: 50) wavefront steps of the pixels two columns ahead of / behind the start
    [REF :165-192]; two rounds of 'extend beyond the track by a 10-px maximum filter, blank the far side of the finish
    line, Gaussian sigma 3, recombine' and one Gaussian sigma 5 of the distance padded with its maximum [REF :236-262];
    distance + 0.08 blurred + 0.06 border-blurred when the map's `use_blurred_factor` is on [REF :270-271]."""
    import yaml
    from scipy import ndimage

    yaml_path = Path(yaml_path)
    with open(yaml_path) as f:
        props = yaml.safe_load(f)
    res = float(props["resolution"])
    ox, oy = float(props["origin"][0]), float(props["origin"][1])
    gray = _read_gray(yaml_path.parent / props["image"])
    H, W = gray.shape
    binary = (gray / np.amax(gray)) > float(props["occupied_thresh"])
    if reference_quirks and H > REFERENCE_CLEARED_PIXEL[0] and W > REFERENCE_CLEARED_PIXEL[1]:
        binary[REFERENCE_CLEARED_PIXEL] = False
    if erosion > 0:   # skimage.morphology.binary_erosion(selem=disk(erosion)): outside the image counts as free
        k = np.arange(-erosion, erosion + 1) ** 2
        binary = ndimage.binary_erosion(binary, structure=np.add.outer(k, k) <= erosion * erosion, border_value=True)
    if use_blurred_factor is None:
        use_blurred_factor = yaml_path.stem not in _NO_BLUR_MAPS
    gx = int((start_xy[0] - ox) / res)
    gy = int(H - (start_xy[1] - oy) / res - 1)

    free = binary.copy()
    finish = np.zeros_like(free)
    for r, step in ((gy, 1), (gy - 1, -1)):          # backward direction: the column one AHEAD of the start
        while free[r, gx + 1]:
            free[r, gx + 1] = False
            finish[r, gx + 1] = True
            r += step

    def area(seed, steps):                           # the seed pixel itself + free cells within `steps` wavefront steps
        d = _wavefront(free, seed, max_dist=steps)
        return d >= 0

    start_big = area((gy, gx + 2), 100) ^ finish
    start_small = area((gy, gx + 2), 50) ^ finish
    target_big = area((gy, gx - 2), 100)

    d = _wavefront(free, (gy, gx))
    reached = d >= 0
    dist = np.where(reached, d, 0).astype(np.float64)
    dist[finish] = float(int(d.max()) + 1)           # loop-exit value of the generator's counter
    dist = dist * res                                # the generator's distance transform already returns metres [REF :221]
    top = np.amax(dist)
    drivable = reached | finish
    drv = drivable.astype(np.float64)
    tgt, st_s, st_b = target_big.astype(np.float64), start_small.astype(np.float64), start_big.astype(np.float64)

    def split_blur(field, sigma):
        """Gaussian of `field` computed twice -- once with the target side of the finish line blanked to the maximum, once
        with the start side blanked to zero -- so that values do not leak across the line; recombined by side."""
        a = top * tgt + field * (1.0 - tgt)
        b = 0.0 * st_s + field * (1.0 - st_s)
        ga = ndimage.gaussian_filter(a, sigma=sigma) * drv
        gb = ndimage.gaussian_filter(b, sigma=sigma) * drv
        return ga * st_b + gb * (1.0 - st_b)

    blurred = dist
    for _ in range(2):
        extended = blurred * drv + ndimage.maximum_filter(blurred, size=10) * (1.0 - drv)
        blurred = split_blur(extended, 3)
    border = split_blur(dist * drv + top * (1.0 - drv), 5)
    out = dist + blurred * 0.08 + border * 0.06 if use_blurred_factor else dist
    out = out * res                                  # ... and the smoothed variant scales once more [REF :273]; it cancels below
    return {"drivable_area": drivable, "norm_distance_to": out / np.amax(out), "distance_to": out,
            "grid_start": (gy, gx)}


def compile_raceline(yaml_path: os.PathLike, start_xy=(0.0, 0.0), reference_quirks: bool = False) -> Dict[str, np.ndarray]:
    """The race-line layer of the reference's map generator [REF docs/maps/costmaps/generate-costmap.py:280-360
    compute_raceline, called at :378 on the ERODED track]: SURVEY.md §8-f4.  The generator only exports it as an image
    (`.spline_line.colorized.png`), it is not one of the keys of `maps.npz` and nothing on the env path reads it; built
    for tools that want the same cost layer.  Returns ``{'raceline'`` (full image, float64, normalised to 1),
    ``'control_points'`` (row, col), ``'spline'`` (1000 samples, row, col)``}``.

    Restated steps: on the track eroded by the map's disk, follow the steepest descent of the smoothed distance to the
    target from ten pixels ahead of the start (8-neighbourhood, rows before columns, first minimum wins) until the path
    meets itself; every `sample_every`-th position is a control point of a closed B-spline of the map's degree, sampled
    1000 times and rasterised; ten Gaussian blurs (sigma 3) of 150 x that line, each masked by the UN-eroded drivable
    area; normalised by its maximum."""
    from scipy import interpolate, ndimage

    yaml_path = Path(yaml_path)
    cfg = generator_settings(yaml_path.stem)
    full = compile_distance_to_target(yaml_path, start_xy, reference_quirks, use_blurred_factor=cfg["blurred"])
    eroded = compile_distance_to_target(yaml_path, start_xy, reference_quirks, use_blurred_factor=cfg["blurred"],
                                        erosion=int(cfg["erosion"]))
    field, free = eroded["distance_to"], eroded["drivable_area"]
    gy, gx = eroded["grid_start"]
    H, W = free.shape
    on_line = np.zeros((H, W), dtype=bool)
    pos = (gy, gx + 10)
    on_line[pos] = True
    cv = [pos]
    steps = ((0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (-1, 1), (1, -1), (-1, -1))   # (row, col), the generator's order
    n = 0
    while True:
        n += 1
        best_val, best = 100000.0, pos
        for dr, dc in steps:
            r, c = pos[0] + dr, pos[1] + dc
            if field[r, c] < best_val and free[r, c]:
                best_val, best = field[r, c], (r, c)
        if n % int(cfg["sample_every"]) == 0:
            cv.append(pos)
        pos = best
        if on_line[pos]:          # the path met itself (or there was nowhere to go)
            break
        on_line[pos] = True
        if n >= 5000:
            break
    cv = np.asarray(cv, dtype=np.int64)
    # closed B-spline through the control polygon [REF :430-456 scipy_bspline(periodic=True)]
    degree, count = int(cfg["degree"]), cv.shape[0]
    kv = np.arange(-degree, count + degree + 1)
    factor, fraction = divmod(count + degree + 1, count)
    ring = np.roll(np.concatenate((cv,) * factor + (cv[:fraction],)), -1, axis=0)
    spline = interpolate.BSpline(kv, ring, degree)(np.linspace(0, count, 1000))
    line = np.zeros((H, W), dtype=np.float64)
    idx = spline.astype(int)
    line[idx[:, 0], idx[:, 1]] = 1.0
    area = full["drivable_area"].astype(np.float64)
    blurred = ndimage.gaussian_filter(line * 150, sigma=3) * area
    for _ in range(9):
        blurred = ndimage.gaussian_filter(blurred, sigma=3) * area
    return {"raceline": blurred / np.amax(blurred), "control_points": cv, "spline": spline}


def _heading_field(tm: TrackMap, rows: np.ndarray, cols: np.ndarray, k: int = 6) -> np.ndarray:
    """Yaw of increasing progress at crop cells (rows, cols): gradient of the wavefront distance over a
    (2k+1)^2 window, using only drivable cells whose distance is within the window's reach (so the
    wrap at the finish line does not leak in).  World frame: +x = +col, +y = -row."""
    d = tm.dist.astype(np.float64)
    drv = tm.drivable
    h, w = drv.shape
    yaw = np.zeros(rows.size)
    for i, (r, c) in enumerate(zip(rows, cols)):
        ra, rb = max(r - k, 0), min(r + k + 1, h)
        ca, cb = max(c - k, 0), min(c + k + 1, w)
        win = d[ra:rb, ca:cb] - d[r, c]
        m = drv[ra:rb, ca:cb] & (np.abs(win) <= 2 * k)
        rr, cc = np.mgrid[ra:rb, ca:cb]
        gx = np.sum(np.where(m, win * (cc - c), 0.0))
        gy = np.sum(np.where(m, win * -(rr - r), 0.0))
        yaw[i] = math.atan2(gy, gx)
    return yaw


def _cell_centre(tm: TrackMap, r: np.ndarray, c: np.ndarray):
    x = tm.origin[0] + (tm.c0 + c + 0.5) * tm.resolution
    y = tm.origin[1] + (tm.full_shape[0] - 1 - (tm.r0 + r) + 0.5) * tm.resolution
    return x, y


def _grid_poses(tm: TrackMap, start_xy) -> np.ndarray:
    """'grid' reset: the start position itself, heading along increasing progress, plus three staggered
    slots ahead of it for multi-car worlds [NEW-SPEC]."""
    row, col = tm.to_pixel(float(start_xy[0]), float(start_xy[1]))
    r, c = row - tm.r0, col - tm.c0
    yaw0 = float(_heading_field(tm, np.array([r]), np.array([c]))[0])
    poses = [(float(start_xy[0]), float(start_xy[1]), yaw0)]
    for k in range(1, 4):
        dx, dy = 0.8 * k, (0.25 if k % 2 else -0.25)
        x = start_xy[0] + dx * math.cos(yaw0) - dy * math.sin(yaw0)
        y = start_xy[1] + dx * math.sin(yaw0) + dy * math.cos(yaw0)
        poses.append((x, y, yaw0))
    return np.asarray(poses, dtype=np.float64)


def _random_reset_poses(tm: TrackMap, clearance_m: float, max_n: int) -> np.ndarray:
    """'random' reset candidates: cell centres with obstacle clearance >= clearance_m, evenly
    subsampled along progress, heading along increasing progress [NEW-SPEC]."""
    need = (clearance_m / tm.resolution) ** 2
    ok = tm.drivable & (tm.edt_sq >= need) & (tm.dist > 12) & (tm.dist < tm.dmax - 12)
    rows, cols = np.nonzero(ok)
    if rows.size == 0:
        return tm.start_poses[:1].copy()
    order = np.lexsort((cols, rows, tm.dist[rows, cols]))
    if order.size > max_n:
        order = order[np.linspace(0, order.size - 1, max_n).astype(np.int64)]
    rows, cols = rows[order], cols[order]
    x, y = _cell_centre(tm, rows, cols)
    yaw = _heading_field(tm, rows, cols)
    return np.stack([x, y, yaw], axis=1).astype(np.float64)


# --------------------------------------------------------------------------------------------
# registry
# --------------------------------------------------------------------------------------------
_cache: Dict[str, TrackMap] = {}


def load_track(name: str) -> TrackMap:
    """Compiled track by racecar_gym name (``austria``) or by map file stem (``f1_aut``)."""
    if name in _cache:
        return _cache[name]
    stem = TRACK_FILES.get(name, name)
    path = DATA_DIR / f"{stem}.npz"
    if not path.exists():
        raise FileNotFoundError(
            f"no compiled track '{name}' ({path}); build it with tools/compile_tracks.py from a map_server yaml")
    tm = TrackMap.load(path)
    tm.name = name
    _cache[name] = tm
    return tm


def available_tracks():
    stems = {p.stem for p in DATA_DIR.glob("*.npz")}
    return sorted([n for n, s in TRACK_FILES.items() if s in stems] + sorted(stems))
