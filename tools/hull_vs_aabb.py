#!/usr/bin/env python3
"""How much fatter than Bullet's collider is ours?  The reference's gripper links collide as the convex hulls of
urdf/meshes/collision/{hand,finger}.obj (Bullet builds a btConvexHullShape from an .obj mesh); the kernels and the oracle use the
AABB of the same vertices as an oriented box (tools/bake_model.py).  Prints volume ratios and the hull's cross-section
against the AABB face at several depths from the two z faces (z = the finger axis of the hand frame: zmax is the finger side).
usage: hull_vs_aabb.py [reference root]     (needs scipy; reads the reference's meshes, so it runs in the build container only)"""
import glob
import sys

import numpy as np
from scipy.spatial import ConvexHull

root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
rng = np.random.default_rng(0)
for name in ("hand", "finger"):
    f = glob.glob(f"{root}/**/collision/{name}.obj", recursive=True)[0]
    V = np.array([[float(x) for x in l.split()[1:4]] for l in open(f) if l.startswith("v ")])
    lo, hi = V.min(0), V.max(0)
    h = ConvexHull(V)
    print(f"{name}.obj: {len(V)} vertices, AABB {np.round(hi - lo, 4)} m, AABB volume {np.prod(hi - lo):.3e} m^3, hull volume {h.volume:.3e} m^3, hull / AABB = {h.volume / np.prod(hi - lo):.3f}")
    for dz in (0.0005, 0.005, 0.01, 0.02, 0.045):
        row = []
        for side in ("zmax", "zmin"):
            z = hi[2] - dz if side == "zmax" else lo[2] + dz
            P = np.column_stack([lo[0] + (hi[0] - lo[0]) * rng.random(40000), lo[1] + (hi[1] - lo[1]) * rng.random(40000), np.full(40000, z)])
            ins = P[np.all(P @ h.equations[:, :3].T + h.equations[:, 3] <= 1e-12, axis=1)]
            ext = f"x [{ins[:, 0].min():+.3f}, {ins[:, 0].max():+.3f}] y [{ins[:, 1].min():+.3f}, {ins[:, 1].max():+.3f}]" if len(ins) else "-"
            row.append(f"{side} - {dz * 1e3:4.1f} mm: section {len(ins) / 40000:.2f} of the AABB face, {ext}")
        print("   " + " | ".join(row))
