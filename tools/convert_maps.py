#!/usr/bin/env python
"""Rasterise the reference's map files into this repo's compact .tpm format (data/maps/).

The GPU box has no /root/reference, so the bench and the `-m gpu` tests read these derived rasters
instead of the reference's .pcd / .bt files.  A .tpm holds only RAW occupancy + the known mask
(bit-packed, run-length coded); inflation is applied by the loader per the occMap contract.

  square_static_map.pcd -> square_static.tpm   400x400x30 @ 0.1 m (occupancy_map.yaml geometry)
  maze.bt, tunnel.bt, box.bt, field.bt -> *.tpm  at the tree's native resolution, occupied bbox + 1 m

Run in the build container:  python tools/convert_maps.py [/root/reference/map]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import trajectory_planner_b200 as tp  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/map"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "maps")
os.makedirs(dst, exist_ok=True)

m = tp.OccMap.from_pcd(os.path.join(src, "square_static_map.pcd"))
m.save_tpm(os.path.join(dst, "square_static.tpm"))
print("square_static", m.info())
for name in ("maze", "tunnel", "box", "field"):
    m = tp.OccMap.from_bt(os.path.join(src, name + ".bt"))
    m.save_tpm(os.path.join(dst, name + ".tpm"))
    print(name, m.info())
