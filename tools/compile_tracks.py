#!/usr/bin/env python
"""Compile ROS map_server maps (yaml + png/pgm) into racing_dreamer_b200/data/tracks/<stem>.npz.

usage: python tools/compile_tracks.py [--maps-dir /root/reference/docs/maps/maps] [stems ...]
The input images are the reference's track data (docs/maps/maps); only the compiled grids are stored.
"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from racing_dreamer_b200 import maps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maps-dir", default="/root/reference/docs/maps/maps")
    ap.add_argument("stems", nargs="*", default=sorted(set(maps.TRACK_FILES.values())))
    args = ap.parse_args()
    maps.DATA_DIR.mkdir(parents=True, exist_ok=True)
    for stem in args.stems:
        t0 = time.time()
        tm = maps.compile_track(Path(args.maps_dir) / f"{stem}.yaml")
        out = maps.DATA_DIR / f"{stem}.npz"
        tm.save(out)
        print(f"{stem}: crop {tm.h}x{tm.w} at (r{tm.r0},c{tm.c0}) lap {tm.lap_length_m():.2f} m "
              f"dmax {tm.dmax} drivable {int(tm.drivable.sum())} cells, {tm.reset_poses.shape[0]} reset poses, "
              f"bits {tm.packed_bits_yup().nbytes} B, start yaw {tm.start_poses[0,2]:+.3f} "
              f"-> {out.name} {out.stat().st_size} B  ({time.time()-t0:.1f}s)")


if __name__ == "__main__":
    main()
