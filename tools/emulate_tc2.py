"""Host-side functional model of render_tc2_kernel (splat_forward_tc.cu): per 4x4x8 tile, Phase A entry words, the
column-per-thread W evaluation with the column / z bit masks, the W x S contraction and the row -> voxel mapping of
the epilogue, in numpy float64, compared against the oracle on a small grid.  Catches mask / mapping slips before the
kernel's first hardware run."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h  # noqa: E402
import oracle  # noqa: E402

kw, inp, variant = h.splat_case("tiny", 3, False, dict(dims=(16, 8, 16), pc_min=(-4.0, -2.0, -4.0)))
a, pi, mi, radii, cov6, dims = h.oracle_prep(kw, inp, variant)
H, W, D = dims
ref = h.oracle_forward(kw, inp, variant)["logits"]
G, C = a["sem"].shape
lo = np.maximum(mi - radii[:, None], 0)
hi = np.minimum(mi + radii[:, None], np.array([H - 1, W - 1, D - 1]))
LOG2E = 1.4426950408889634
coef = np.stack([-0.5 * LOG2E * cov6[:, 0], -0.5 * LOG2E * cov6[:, 1], -0.5 * LOG2E * cov6[:, 2],
                 -LOG2E * cov6[:, 3], -LOG2E * cov6[:, 4], -LOG2E * cov6[:, 5]], 1).astype(np.float64)
S = (a["opa"][:, None] * a["sem"]).astype(np.float64)            # base variant: opacity folded into the class vector
pts = a["pts"].reshape(H, W, D, 3).astype(np.float64)
out = np.zeros((H, W, D, C))
for bx in range(0, H, 4):
    for by in range(0, W, 4):
        for bz in range(0, D, 8):
            # Phase A: ascending Gaussians whose box meets the tile, entry = x mask | y mask << 4 | z mask << 8
            entries = []
            for g in range(G):
                if lo[g, 0] <= bx + 3 and hi[g, 0] >= bx and lo[g, 1] <= by + 3 and hi[g, 1] >= by and lo[g, 2] <= bz + 7 and hi[g, 2] >= bz:
                    rx0, rx1 = max(lo[g, 0] - bx, 0), min(hi[g, 0] - bx, 3)
                    ry0, ry1 = max(lo[g, 1] - by, 0), min(hi[g, 1] - by, 3)
                    rz0, rz1 = max(lo[g, 2] - bz, 0), min(hi[g, 2] - bz, 7)
                    xm = ((2 << rx1) - 1) & ~((1 << rx0) - 1)
                    ym = ((2 << ry1) - 1) & ~((1 << ry0) - 1)
                    zm = ((2 << rz1) - 1) & ~((1 << rz0) - 1)
                    entries.append((xm | (ym << 4) | (zm << 8), g))
            Wm = np.zeros((128, len(entries)))
            for pcol in range(16):                       # producer: column pcol = 4*cy + cx
                cx, cy = pcol & 3, pcol >> 2
                col_bits = (1 << cx) | (1 << (4 + cy))
                cpx, cpy = pts[bx + cx, by + cy, bz, 0], pts[bx + cx, by + cy, bz, 1]
                for j, (ent, g) in enumerate(entries):
                    zm = (ent >> 8) & 0xff if (ent & col_bits) == col_bits else 0
                    dx, dy = a["means"][g, 0] - cpx, a["means"][g, 1] - cpy
                    A = (coef[g, 0] * dx + coef[g, 3] * dy) * dx + coef[g, 1] * dy * dy
                    B = coef[g, 4] * dy + coef[g, 5] * dx
                    for z in range(8):
                        dz = a["means"][g, 2] - pts[bx + cx, by + cy, bz + z, 2]
                        q = (coef[g, 2] * dz + B) * dz + A
                        Wm[16 * z + pcol, j] = 2.0 ** q if (zm >> z) & 1 else 0.0
            acc = Wm @ S[[g for _, g in entries]] if entries else np.zeros((128, C))
            for r in range(128):                         # epilogue: row r = 16*z + column
                ez, ecol = r >> 4, r & 15
                out[bx + (ecol & 3), by + (ecol >> 2), bz + ez] = acc[r]
err = np.abs(out.reshape(-1, C) - ref)
tol = h.ATOL + h.RTOL * np.abs(ref)
print("max abs err", err.max(), "violations", int((err > tol).sum()), "of", err.size)
assert not (err > tol).any()
print("tc2 model matches the oracle")
