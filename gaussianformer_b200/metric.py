"""mIoU / occupancy-IoU exactly as the reference accumulates them (``misc/metric_util.py:9-111``):
per-class seen / correct / positive counters (+ one "non-empty" slot), three ``all_reduce`` calls at
epoch end, classes never seen count as IoU 1.  Works on any device (counters live where the
predictions live) and with any ``torch.distributed`` backend."""
from __future__ import annotations

import torch
import torch.distributed as dist


class MeanIoU:
    def __init__(self, class_indices, empty_label, use_mask=False):
        self.class_indices = list(class_indices)
        self.num_classes = len(self.class_indices)
        self.empty_label = empty_label
        self.use_mask = use_mask
        self.reset()

    def reset(self, device="cpu"):
        self.total_seen = torch.zeros(self.num_classes + 1, dtype=torch.float64, device=device)
        self.total_correct = torch.zeros_like(self.total_seen)
        self.total_positive = torch.zeros_like(self.total_seen)

    def after_step(self, outputs, targets, mask=None):
        if self.total_seen.device != outputs.device:   # keep what was accumulated: move the counters, never wipe them
            self.total_seen = self.total_seen.to(outputs.device)
            self.total_correct = self.total_correct.to(outputs.device)
            self.total_positive = self.total_positive.to(outputs.device)
        if mask is not None:
            outputs, targets = outputs[mask], targets[mask]
        for i, c in enumerate(self.class_indices):
            self.total_seen[i] += (targets == c).sum()
            self.total_correct[i] += ((targets == c) & (outputs == c)).sum()
            self.total_positive[i] += (outputs == c).sum()
        occ_t, occ_o = targets != self.empty_label, outputs != self.empty_label
        self.total_seen[-1] += occ_t.sum()
        self.total_correct[-1] += (occ_t & occ_o).sum()
        self.total_positive[-1] += occ_o.sum()

    def after_epoch(self):
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.total_seen)
            dist.all_reduce(self.total_correct)
            dist.all_reduce(self.total_positive)
        seen, correct, positive = (t.cpu() for t in (self.total_seen, self.total_correct, self.total_positive))
        ious = []
        for i in range(self.num_classes):
            if seen[i] == 0:
                ious.append(1.0)
            else:
                ious.append(float(correct[i] / (seen[i] + positive[i] - correct[i])))
        occ = float(correct[-1] / (seen[-1] + positive[-1] - correct[-1]))
        return sum(ious) / len(ious) * 100.0, occ * 100.0


def synthetic_labels(oracle_logits, num_classes, flip_fraction=0.1, seed=1):
    """labels = arg-max of the oracle's logits with ``flip_fraction`` of the voxels re-drawn
    uniformly from ``0..num_classes-1`` (seeded); mask = ``label != 0`` (dataset/transform_3d.py:509)."""
    logits = torch.as_tensor(oracle_logits)
    labels = logits.argmax(dim=1)
    gen = torch.Generator().manual_seed(seed)
    flip = torch.rand(labels.shape[0], generator=gen) < flip_fraction
    labels = torch.where(flip, torch.randint(0, num_classes, labels.shape, generator=gen), labels)
    return labels, labels != 0


def miou_parity(pred_new, pred_ref, labels, mask, num_classes, empty_label=None):
    """mIoU / occupancy IoU of two arg-max predictions against the same labels, as the reference evaluates them
    (``misc/metric_util.py:35-111``: classes 1..C-2 scored, class C-1 = empty).  Returns
    ``{"new": (miou, iou), "ref": (miou, iou), "abs_diff": max |difference| in mIoU points}``."""
    empty_label = num_classes - 1 if empty_label is None else empty_label
    out = {}
    for name, pred in (("new", pred_new), ("ref", pred_ref)):
        m = MeanIoU(list(range(1, num_classes - 1)), empty_label=empty_label)
        m.after_step(torch.as_tensor(pred).long().cpu(), labels, mask)
        out[name] = m.after_epoch()
    out["abs_diff"] = max(abs(out["new"][0] - out["ref"][0]), abs(out["new"][1] - out["ref"][1]))
    return out
