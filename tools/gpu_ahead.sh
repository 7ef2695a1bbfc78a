#!/bin/bash
# draw-ahead on/off (RD_LIDAR_AHEAD) with the longest-first order, small and medium launches
for spec in austria:4096:1 austria:1024:1 columbia:16384:1 treitlstrasse_v2:8192:1; do
  for rep in 1 2; do for o in 0 1; do
    echo -n "ahead $o "; RD_LIDAR_AHEAD=$o RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | cut -c1-130
  done; done
done
