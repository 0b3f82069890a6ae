"""Running averages and top-k accuracy, interface of the reference's utils/meters.py:4-20,59-72."""
import torch


class AverageMeter(object):
    """Tracks the latest value and the count-weighted mean of a scalar series."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count else 0


def accuracy(output, target, topk=(1,)):
    """precision@k in percent for each k, as 0-dim tensors (same contract as utils/meters.py:59-72)."""
    with torch.no_grad():
        kmax = min(max(topk), output.size(1))
        ranked = output.topk(kmax, dim=1, largest=True, sorted=True).indices  # [B, kmax]
        hits = ranked.eq(target.view(-1, 1).to(ranked.dtype))
        scale = 100.0 / target.size(0)
        return [hits[:, :min(k, kmax)].any(dim=1).float().sum() * scale for k in topk]


class AccuracyMeter(object):
    def __init__(self, topk=(1,)):
        self.topk = topk
        self.reset()

    def reset(self):
        self._meters = {k: AverageMeter() for k in self.topk}

    def update(self, output, target):
        for k, v in zip(self.topk, accuracy(output, target, self.topk)):
            self._meters[k].update(float(v))

    @property
    def val(self):
        return {k: m.val for k, m in self._meters.items()}

    @property
    def avg(self):
        return {k: m.avg for k, m in self._meters.items()}

    @property
    def avg_error(self):
        return {k: 100.0 - m.avg for k, m in self._meters.items()}
