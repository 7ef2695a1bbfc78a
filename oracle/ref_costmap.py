"""Run the UNMODIFIED reference map generator in the build container (test infrastructure).

``docs/maps/costmaps/generate-costmap.py`` [REF] imports ``skimage`` and ``cmapy`` (absent here) and uses the removed
NumPy aliases ``np.float`` / ``np.bool`` / ``np.int``.  ``load_generator()`` installs stand-ins for exactly the names
the script touches -- ``skimage.io.imread(as_gray=True)`` (Pillow + the rgb2gray weights skimage documents),
``skimage.morphology.binary_dilation / binary_erosion / selem.square / selem.disk`` (thin wrappers over
``scipy.ndimage``, which is what skimage itself calls), ``img_as_ubyte`` and a no-op ``cmapy`` -- and loads the script
from ``/root/reference`` (nothing is copied).  Used by tests that are skipped when /root/reference is absent.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io as _io
import sys
import types
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path("/root/reference")
SCRIPT = REFERENCE_ROOT / "docs" / "maps" / "costmaps" / "generate-costmap.py"


def _imread(path, as_gray=False):
    from racing_dreamer_b200.maps import _read_gray   # the reader under test is the product's own
    assert as_gray
    return _read_gray(Path(path))


def load_generator():
    from scipy import ndimage
    if not SCRIPT.exists():
        raise FileNotFoundError(SCRIPT)
    selem = types.SimpleNamespace(
        square=lambda width, dtype=bool: np.ones((width, width), dtype=dtype),
        disk=lambda radius, dtype=bool: (np.add.outer(np.arange(-radius, radius + 1) ** 2,
                                                       np.arange(-radius, radius + 1) ** 2) <= radius * radius).astype(dtype))
    morphology = types.ModuleType("skimage.morphology")
    morphology.selem = selem

    def binary_dilation(image, selem=None):
        """ndimage.binary_dilation restricted to the bounding box of the set pixels (+ the structure's reach): same
        result, far less work for the generator's one-pixel-per-iteration wavefronts."""
        image = np.asarray(image, dtype=bool)
        rows, cols = np.flatnonzero(image.any(axis=1)), np.flatnonzero(image.any(axis=0))
        out = np.zeros_like(image)
        if rows.size == 0:
            return out
        k = max(selem.shape) if selem is not None else 3
        r0, r1 = max(rows[0] - k, 0), min(rows[-1] + k + 1, image.shape[0])
        c0, c1 = max(cols[0] - k, 0), min(cols[-1] + k + 1, image.shape[1])
        out[r0:r1, c0:c1] = ndimage.binary_dilation(image[r0:r1, c0:c1], structure=selem)
        return out

    morphology.binary_dilation = binary_dilation
    morphology.binary_erosion = lambda image, selem=None: ndimage.binary_erosion(image, structure=selem, border_value=True)
    skio = types.ModuleType("skimage.io")
    skio.imread = _imread
    skio.imsave = lambda *a, **k: None
    skimage = types.ModuleType("skimage")
    skimage.io, skimage.morphology = skio, morphology
    skimage.img_as_ubyte = lambda a: (np.clip(np.asarray(a, dtype=np.float64), 0, 1) * 255).astype(np.uint8)
    cmapy = types.ModuleType("cmapy")
    cmapy.colorize = lambda img, *a, **k: np.repeat(np.asarray(img)[..., None], 3, axis=2)
    mods = {"skimage": skimage, "skimage.io": skio, "skimage.morphology": morphology, "cmapy": cmapy}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    for name, typ in (("float", float), ("bool", bool), ("int", int)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    try:
        spec = importlib.util.spec_from_file_location("ref_generate_costmap", SCRIPT)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def reference_layers(yaml_path, start_xy=(0.0, 0.0), full_run=False, out_path=None):
    """-> dict(drivable_area, norm_distance_from_start, norm_distance_to_obstacle[, norm_distance_to]) as the reference
    computes them.  full_run=False calls the forward distance transform and the three EDT statements of ``run()``
    [REF generate-costmap.py:365-382]; full_run=True calls ``run()`` itself (adds the smoothed distance-to-target; slow)."""
    from scipy import ndimage
    mod = load_generator()
    with contextlib.redirect_stdout(_io.StringIO()):
        gen = mod.CostmapGenerator(starting_position=start_xy, input_yaml_path=str(yaml_path),
                                   output_path=str(out_path or "/tmp/ref_costmap_out.npz"))
        gen.verbose = False
        if full_run:
            gen.run()
            z = np.load(gen.output_path)
            return {k: z[k] for k in z.files}
        drivable, _, norm, *_ = gen.compute_distance_transform(starting_position=gen.grid_starting_position,
                                                               forward_direction=True)
        d = ndimage.distance_transform_edt(drivable, return_distances=True, return_indices=False)
        d = d.astype(float) * float(gen.map_properties["resolution"])
    return {"drivable_area": drivable, "norm_distance_from_start": norm,
            "norm_distance_to_obstacle": d / np.amax(d.flatten()),
            "grid_starting_position": np.asarray(gen.grid_starting_position)}


def reference_raceline(yaml_path, start_xy=(0.0, 0.0), out_path=None):
    """-> (race-line layer of the ERODED track as the reference's run() computes it [REF generate-costmap.py:378,
    compute_raceline :280-360], the generator's per-map settings).  run() only exports that layer as an image, so the
    UNMODIFIED method is wrapped to keep what it returns; nothing of the reference is changed."""
    mod = load_generator()
    with contextlib.redirect_stdout(_io.StringIO()):
        gen = mod.CostmapGenerator(starting_position=start_xy, input_yaml_path=str(yaml_path),
                                   output_path=str(out_path or "/tmp/ref_raceline_out.npz"))
        gen.verbose = False
        kept = {}
        original = gen.compute_raceline

        def keep(*args):
            kept[args[4]] = original(*args)
            return kept[args[4]]

        gen.compute_raceline = keep
        gen.run()
    settings = dict(erosion=gen.erosion_value, degree=gen.d_value, blurred=gen.use_blurred_factor, sample_every=gen.sample_every)
    return kept["eroded"], settings

