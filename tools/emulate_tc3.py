"""Host-side functional model of render_tc3_kernel (splat_forward_tc.cu): the SIMT bin (8 x 4 columns x 16 z) split
into four 4 x 4 x 8 tiles, per-tile compaction of every 32-entry batch by the tile's x / z bit ranges, the
column-per-lane W evaluation with the column / z masks, K positions in steps of 8 with zero padding, W x S, and the
row -> voxel mapping of the epilogue; numpy float64 against the oracle on a small grid."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h  # noqa: E402

kw, inp, variant = h.splat_case("tiny", 3, False, dict(dims=(16, 8, 32), pc_min=(-4.0, -2.0, -8.0)))
a, pi, mi, radii, cov6, dims = h.oracle_prep(kw, inp, variant)
H, W, D = dims
ref = h.oracle_forward(kw, inp, variant)["logits"]
G, C = a["sem"].shape
lo = np.maximum(mi - radii[:, None], 0)
hi = np.minimum(mi + radii[:, None], np.array([H - 1, W - 1, D - 1]))
LOG2E = 1.4426950408889634
coef = np.stack([-0.5 * LOG2E * cov6[:, 0], -0.5 * LOG2E * cov6[:, 1], -0.5 * LOG2E * cov6[:, 2],
                 -LOG2E * cov6[:, 3], -LOG2E * cov6[:, 4], -LOG2E * cov6[:, 5]], 1).astype(np.float64)
S = (a["opa"][:, None] * a["sem"]).astype(np.float64)
pts = a["pts"].reshape(H, W, D, 3).astype(np.float64)
out = np.zeros((H, W, D, C))
steps = 0
for bx in range(0, H, 8):
    for by in range(0, W, 4):
        for bz in range(0, D, 16):
            entries = []
            for g in range(G):
                if lo[g, 0] <= bx + 7 and hi[g, 0] >= bx and lo[g, 1] <= by + 3 and hi[g, 1] >= by and lo[g, 2] <= bz + 15 and hi[g, 2] >= bz:
                    rx0, rx1 = max(lo[g, 0] - bx, 0), min(hi[g, 0] - bx, 7)
                    ry0, ry1 = max(lo[g, 1] - by, 0), min(hi[g, 1] - by, 3)
                    rz0, rz1 = max(lo[g, 2] - bz, 0), min(hi[g, 2] - bz, 15)
                    xm = ((2 << rx1) - 1) & ~((1 << rx0) - 1)
                    ym = ((2 << ry1) - 1) & ~((1 << ry0) - 1)
                    zm = ((2 << rz1) - 1) & ~((1 << rz0) - 1)
                    entries.append((int(xm) | (int(ym) << 8) | (int(zm) << 16), g))
            for warp in range(4):                                   # one tile per warp
                tile_x, tile_z = 0xF << (4 * (warp & 1)), 0xFF << (16 + 8 * (warp >> 1))
                zshift = 16 + 8 * (warp >> 1)
                acc = np.zeros((128, C))
                for b0 in range(0, len(entries), 32):               # batches of 32 entries
                    batch = entries[b0:b0 + 32]
                    hits = [j for j, (ent, g) in enumerate(batch) if (ent & tile_x) and (ent & tile_z)]
                    for base in range(0, len(hits), 8):             # a step fills 8 K positions
                        steps += 1
                        Wm = np.zeros((128, 8))
                        Sm = np.zeros((8, C))
                        for q in range(8):
                            if base + q >= len(hits):
                                continue                            # zero padding
                            ent, g = batch[hits[base + q]]
                            Sm[q] = S[g]
                            for pcol in range(16):
                                col_x, col_y = 1 << (4 * (warp & 1) + (pcol & 3)), 1 << (8 + (pcol >> 2))
                                zm = (ent >> zshift) & 0xff if (ent & col_x) and (ent & col_y) else 0
                                X, Y = bx + 4 * (warp & 1) + (pcol & 3), by + (pcol >> 2)
                                cpx, cpy = pts[X, Y, bz + 8 * (warp >> 1), 0], pts[X, Y, bz + 8 * (warp >> 1), 1]
                                dx, dy = a["means"][g, 0] - cpx, a["means"][g, 1] - cpy
                                A = (coef[g, 0] * dx + coef[g, 3] * dy) * dx + coef[g, 1] * dy * dy
                                B = coef[g, 4] * dy + coef[g, 5] * dx
                                for z in range(8):
                                    dz = a["means"][g, 2] - pts[X, Y, bz + 8 * (warp >> 1) + z, 2]
                                    Wm[16 * z + pcol, q] = 2.0 ** ((coef[g, 2] * dz + B) * dz + A) if (zm >> z) & 1 else 0.0
                        acc += Wm @ Sm
                for r in range(128):                                # epilogue: row r of tile `warp`
                    rz, rcol = r >> 4, r & 15
                    out[bx + 4 * (warp & 1) + (rcol & 3), by + (rcol >> 2), bz + 8 * (warp >> 1) + rz] = acc[r]
err = np.abs(out.reshape(-1, C) - ref)
tol = h.ATOL + h.RTOL * np.abs(ref)
print("steps", steps, "max abs err", err.max(), "violations", int((err > tol).sum()), "of", err.size)
assert not (err > tol).any()
print("tc3 model matches the oracle")
