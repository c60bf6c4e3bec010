"""Oracle OHEM cross-entropy and accuracy (TEST INFRASTRUCTURE).

Follows ``mmseg/models/losses/ohem_cross_entropy_loss.py:37-90`` and
``mmseg/models/losses/accuracy.py:6-61``.
"""
import torch
import torch.nn.functional as F


def ohem_cross_entropy(score, target, ignore_label=255, thres=0.7, min_kept=100000,
                       loss_weight=1.0, class_weight=None):
    """score [N,K,H,W] float, target [N,H,W] int64 -> scalar.

    ohem_cross_entropy_loss.py:62-90: softmax; per-pixel (class-weighted) CE with
    ignore; probability of the true class; ascending sort of the valid ones;
    ``min_value = sorted[min(min_kept, n-1)]``; ``threshold = max(min_value, thres)``;
    plain mean of CE over valid pixels with ``prob < threshold``; times loss_weight.
    """
    min_kept = max(1, min_kept)                                  # ctor, :47
    prob = F.softmax(score, dim=1)
    w = score.new_tensor(class_weight) if class_weight is not None else None
    pix = F.cross_entropy(score, target, weight=w, ignore_index=ignore_label,
                          reduction='none').contiguous().view(-1)
    mask = target.contiguous().view(-1) != ignore_label
    tmp = target.clone()
    tmp[tmp == ignore_label] = 0
    p = prob.gather(1, tmp.unsqueeze(1)).contiguous().view(-1)[mask].contiguous()
    p, ind = p.sort()
    if p.numel() == 0:
        return score.new_tensor(0.0)
    min_value = p[min(min_kept, p.numel() - 1)]
    threshold = max(min_value, thres)
    pix = pix[mask][ind]
    return loss_weight * pix[p < threshold].mean()


def accuracy(pred, target, ignore_index=None):
    """accuracy.py:41-60 at topk=1, thresh=None: (correct+eps)*100/(total+eps)."""
    if pred.size(0) == 0:
        return pred.new_tensor(0.)
    label = pred.topk(1, dim=1)[1].transpose(0, 1)
    correct = label.eq(target.unsqueeze(0).expand_as(label))
    eps = torch.finfo(torch.float32).eps
    if ignore_index is not None:
        correct = correct[:, target != ignore_index]
        total = target[target != ignore_index].numel() + eps
    else:
        total = target.numel() + eps
    c = correct[:1].reshape(-1).float().sum(0, keepdim=True) + eps
    return c.mul_(100.0 / total)
