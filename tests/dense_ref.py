"""Dense O(N*G) float64 torch formulas of the splat op, used ONLY to validate the oracle on tiny
problems (independent of the oracle's list construction).  Semantics: SURVEY.md §8(a')."""
import math

import torch


def inclusion_mask(points_int, means_int, radii):
    """mask[n,g] = voxel(n) inside Gaussian g's integer box (radii [G] or [G,3])."""
    r = radii if radii.dim() == 2 else radii[:, None].expand(-1, 3)
    lo = (means_int - r)[None]          # 1,G,3
    hi = (means_int + r)[None]
    p = points_int[:, None, :]          # N,1,3
    return ((p >= lo) & (p <= hi)).all(-1)


def power_matrix(pts, means, cov6):
    d = means[None, :, :] - pts[:, None, :]     # N,G,3
    dx, dy, dz = d.unbind(-1)
    a, b, c, dd, e, f = cov6.unbind(-1)
    return -0.5 * (a * dx * dx + b * dy * dy + c * dz * dz) - (dd * dx * dy + e * dy * dz + f * dx * dz)


def base_forward(pts, points_int, means, means_int, opa, sem, cov6, radii):
    mask = inclusion_mask(points_int, means_int, radii).to(pts.dtype)
    w = mask * opa[None] * torch.exp(power_matrix(pts, means, cov6))
    return w @ sem


def prob_forward(pts, points_int, means, means_int, opa, sem, cov6, radii):
    mask = inclusion_mask(points_int, means_int, radii)
    E = torch.exp(power_matrix(pts, means, cov6)) * mask
    a, b, c, d, e, f = cov6.unbind(-1)
    det = a * b * c + 2 * d * e * f - a * e * e - b * f * f - c * d * d
    kappa = (2 * 3.1415926535) ** -1.5
    P = kappa * torch.sqrt(det)[None] * E * opa[None]
    Z = P.sum(1)
    C = sem.shape[1]
    safe = Z > 1e-9
    logits = (P @ sem) / torch.where(safe, Z, torch.ones_like(Z))[:, None]
    fallback = torch.full_like(logits, 1.0 / (C - 1))
    fallback[:, C - 1] = 0
    logits = torch.where(safe[:, None], logits, fallback)
    keep = torch.where(mask, 1 - E, torch.ones_like(E)).prod(1)
    return logits, 1 - keep, E.sum(1), Z
