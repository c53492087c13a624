"""Host-side count of the render kernel's walk steps under different batch pooling schemes (no GPU).

For every bin (8 x 4 columns x 16 z) of config 2's synthetic sample: the ordered list of Gaussians whose clipped box
overlaps the bin; for every warp (4 x 4 columns x 8 z = 16 columns x 2 z quads) the per-lane hit sets.  A warp runs
max-over-lanes steps per pooled group of records.  Prints total warp steps for pools of 32, 64, 96 records and the
whole list, plus the lane and z-quad utilisation.
"""
import sys
import numpy as np

sys.path.insert(0, ".")
from gaussianformer_b200.synthetic import make_splat_inputs  # noqa: E402
from oracle import host_prep  # noqa: E402


def main(name="gs25600_solid"):
    kw, inp, variant = make_splat_inputs(name, seed=0)
    H, W, D = kw["H"], kw["W"], kw["D"]
    pts = inp["pts"][0].numpy(); means = inp["means"][0].numpy(); scales = inp["scales"][0].numpy()
    _, mi, radii = host_prep(pts, means, scales, np.asarray(kw["pc_min"], np.float32), kw["grid_size"],
                             kw["scale_multiplier"])
    dims = np.array([H, W, D])
    lo = np.clip(mi - radii[:, None], 0, dims - 1)
    hi = np.clip(mi + radii[:, None], 0, dims - 1)
    G = len(mi)
    print("G", G, "mean box edge", (hi - lo + 1).mean(0))
    rng = np.random.default_rng(0)
    nbx, nby = H // 8, W // 4
    bins = [(bx, by) for bx in range(nbx) for by in range(nby)]
    sel = rng.choice(len(bins), size=int(sys.argv[2]) if len(sys.argv) > 2 else 300, replace=False)
    tot = {32: 0, 64: 0, 96: 0, "all": 0}
    win = {2: 0, 3: 0, 4: 0}
    rowwin = {4: 0, 8: 0, 16: 0}                     # steps when a step may only touch rows [j0, j0 + R) of the batch
    rows_per_step = {4: 0, 8: 0, 16: 0, 32: 0}
    lane_steps = 0; vox_pairs = 0; nrec = 0; uniform = 0
    for s in sel:
        bx, by = bins[s]
        x0, y0 = bx * 8, by * 4
        m = (hi[:, 0] >= x0) & (lo[:, 0] < x0 + 8) & (hi[:, 1] >= y0) & (lo[:, 1] < y0 + 4)
        idx = np.nonzero(m)[0]
        nrec += len(idx)
        L, Hh = lo[idx], hi[idx]
        for w in range(4):          # warp: x half (4 of 8 columns), z half (8 of 16)
            wx0 = x0 + 4 * (w & 1); wz0 = 8 * (w >> 1)
            hits = np.zeros((len(idx), 32), bool); cover = np.zeros((len(idx), 32), np.int32)
            for lane in range(32):
                cx = wx0 + (lane & 3); cy = y0 + ((lane >> 2) & 3); z0 = wz0 + 4 * (lane >> 4)
                inside = (L[:, 0] <= cx) & (Hh[:, 0] >= cx) & (L[:, 1] <= cy) & (Hh[:, 1] >= cy)
                zc = np.minimum(Hh[:, 2], z0 + 3) - np.maximum(L[:, 2], z0) + 1
                h = inside & (zc > 0)
                hits[:, lane] = h; cover[:, lane] = np.where(h, zc, 0)
            lane_steps += hits.sum(); vox_pairs += cover.sum(); uniform += hits.any(1).sum()
            for pool in (32, 64, 96):
                for b in range(0, len(idx), pool):
                    tot[pool] += hits[b:b + pool].sum(0).max() if len(idx) else 0
            tot["all"] += hits.sum(0).max() if len(idx) else 0
            nb = (len(idx) + 31) // 32
            per = np.zeros((nb, 32), np.int64)           # hits of lane l in batch b
            for b in range(nb):
                per[b] = hits[32 * b:32 * b + 32].sum(0)
            for R in rows_per_step:
                for b in range(nb):
                    pend = [list(np.nonzero(hits[32 * b:32 * b + 32, l])[0]) for l in range(32)]
                    while any(pend):
                        j0 = min(q[0] for q in pend if q)
                        touched = set()
                        for q in pend:
                            if q and q[0] < j0 + R:
                                touched.add(q.pop(0))
                        if R in rowwin:
                            rowwin[R] += 1
                        rows_per_step[R] += len(touched)
                        if R == 32:
                            rowwin.setdefault(32, 0); rowwin[32] += 1
            for Wn in win:
                rem = per.copy()
                for b in range(nb):
                    n = rem[b].max()
                    win[Wn] += n
                    budget = np.full(32, n)
                    for bb in range(b, min(b + Wn, nb)):
                        take = np.minimum(rem[bb], budget)
                        rem[bb] -= take; budget -= take
    nw = 4 * len(sel)
    print("records per bin", nrec / len(sel))
    for k, v in tot.items():
        print("pool", k, "steps per warp", v / nw, "lane utilisation", lane_steps / (32.0 * v))
    for k, v in win.items():
        print("drain-oldest window", k, "steps per warp", v / nw)
    for k, v in rowwin.items():
        print("row window", k, "steps per warp", v / nw, "distinct rows per step", rows_per_step[k] / max(v, 1))
    print("warp-uniform steps per warp", uniform / nw)
    print("z-quad utilisation", vox_pairs / (4.0 * lane_steps))


if __name__ == "__main__":
    main(*sys.argv[1:2])
