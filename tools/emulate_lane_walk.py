"""Host-side model of the render kernel's Phase B on the bench workload: how many steps does a warp take
per bin (a) when all lanes march through every record that touches the warp's footprint (first generation)
and (b) when every lane walks its own hits, per batch of 32 / 64 / all records (lane-private traversal),
for the lane->voxel mappings that were tried.  Pure numpy, no GPU; used to size the change before building it
(DESIGN.md 4.1)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gaussianformer_b200.synthetic import make_splat_inputs  # noqa: E402

kw, inp, _ = make_splat_inputs("gs25600_solid", seed=0)
means, scales = inp["means"][0].numpy(), inp["scales"][0].numpy()
H, W, D = kw["H"], kw["W"], kw["D"]
pc, gs = np.array(kw["pc_min"], np.float32), np.float32(kw["grid_size"])
mi = ((means - pc) / gs).astype(np.int32)
r = np.ceil(scales.max(1) * np.float32(kw["scale_multiplier"]) / gs).astype(np.int32)
lo = np.maximum(mi - r[:, None], 0)
hi = np.minimum(mi + r[:, None], np.array([H - 1, W - 1, D - 1]))
VOX = 4
MAPS = {
    "map0: 4x4 columns x 2 z quads": lambda w, l: ((w & 1) * 4 + (l >> 3), (l >> 1) & 3, (w >> 1) * 2 + (l & 1)),
    "map1: 8x4 columns x 1 z quad": lambda w, l: (l & 7, l >> 3, w),
}
rng = np.random.default_rng(0)
bins = [(rng.integers(0, H // 8) * 8, rng.integers(0, W // 4) * 4) for _ in range(150)]
for name, f in MAPS.items():
    touching = 0
    steps = {32: 0, 64: 0, 1 << 30: 0}
    lane_hits = 0.0
    for bx, by in bins:
        sel = np.where((lo[:, 0] <= bx + 7) & (hi[:, 0] >= bx) & (lo[:, 1] <= by + 3) & (hi[:, 1] >= by))[0]
        l0, h0, n = lo[sel], hi[sel], len(sel)
        for warp in range(4):
            cov = np.zeros((n, 32), bool)
            for lane in range(32):
                lx, ly, lq = f(warp, lane)
                X, Y, z0 = bx + lx, by + ly, VOX * lq
                cov[:, lane] = ((l0[:, 0] <= X) & (h0[:, 0] >= X) & (l0[:, 1] <= Y) & (h0[:, 1] >= Y) &
                                (np.minimum(h0[:, 2], z0 + VOX - 1) >= np.maximum(l0[:, 2], z0)))
            touching += cov.any(1).sum()
            lane_hits += cov.sum() / 32
            for B in steps:
                for s in range(0, n, min(B, max(n, 1))):
                    steps[B] += cov[s:s + B].sum(0).max()
    nw = 4 * len(bins)
    print(f"{name}: records touching the footprint {touching / nw:.1f}/warp; mean hits per lane {lane_hits / nw:.1f}; "
          + "; ".join(f"lane walk, batch {'all' if B > 64 else B}: {v / nw:.1f} steps/warp" for B, v in steps.items()))
