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
        if self.total_seen.device != outputs.device:
            self.reset(outputs.device)
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
