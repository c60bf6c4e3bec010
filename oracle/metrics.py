"""Oracle IoU metric / confusion matrix (TEST INFRASTRUCTURE).

Follows ``mmseg/evaluation/metrics/iou_metric.py:163-200`` (intersect_and_union),
``:102-161`` (compute_metrics), ``:202-295`` (total_area_to_metrics) and
``tools/analysis_tools/confusion_matrix.py:50-77`` (bincount(K*gt+pred), rows = GT).
"""
from collections import OrderedDict

import numpy as np
import torch


def intersect_and_union(pred_label, label, num_classes, ignore_index):
    """iou_metric.py:185-199: three float32 ``histc`` over the non-ignored pixels."""
    mask = label != ignore_index
    pred_label, label = pred_label[mask], label[mask]
    inter = pred_label[pred_label == label]
    k = num_classes
    area_i = torch.histc(inter.float(), bins=k, min=0, max=k - 1)
    area_p = torch.histc(pred_label.float(), bins=k, min=0, max=k - 1)
    area_l = torch.histc(label.float(), bins=k, min=0, max=k - 1)
    return area_i, area_p + area_l - area_i, area_p, area_l


def confusion_matrix(pred, gt, num_classes, ignore_index):
    """confusion_matrix.py:66-74 as int64; GT values outside [0,K) that are not
    ``ignore_index`` land in an extra spill row K (the reference's uint8 bincount
    would mis-bin them; ``histc`` drops them from area_label but still counts
    the pixel in area_pred_label - the spill row reproduces that)."""
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    gt = np.asarray(gt).reshape(-1).astype(np.int64)
    keep = gt != ignore_index
    pred, gt = pred[keep], gt[keep]
    gt = np.where((gt < 0) | (gt >= num_classes), num_classes, gt)
    n = num_classes
    return np.bincount(n * gt + pred, minlength=(n + 1) * n).reshape(n + 1, n)


def confusion_to_areas(cm):
    """[K+1,K] (or [K,K]) confusion matrix -> the four histograms of iou_metric.py."""
    cm = np.asarray(cm, dtype=np.int64)
    k = cm.shape[1]
    inter = np.diag(cm[:k])
    area_l = cm[:k].sum(1)
    area_p = cm.sum(0)
    return inter, area_p + area_l - inter, area_p, area_l


def total_area_to_metrics(total_i, total_u, total_p, total_l, metrics=('mIoU',),
                          nan_to_num=None, beta=1):
    """iou_metric.py:202-295 on torch float tensors (division semantics kept)."""
    if isinstance(metrics, str):
        metrics = [metrics]
    if not set(metrics).issubset({'mIoU', 'mDice', 'mFscore'}):
        raise KeyError(f'metrics {metrics} is not supported')
    t = [torch.as_tensor(np.asarray(a), dtype=torch.float32) if not torch.is_tensor(a)
         else a for a in (total_i, total_u, total_p, total_l)]
    ti, tu, tp, tl = t

    def f_score(p, r):
        return (1 + beta**2) * (p * r) / ((beta**2 * p) + r)

    out = OrderedDict(aAcc=ti.sum() / tl.sum())
    for m in metrics:
        if m == 'mIoU':
            out['IoU'], out['Acc'] = ti / tu, ti / tl
        if m in ('mIoU', 'mFscore'):
            p, r = ti / tp, ti / tl
            out['Fscore'] = torch.tensor([f_score(a, b) for a, b in zip(p, r)])
            out['Precision'], out['Recall'] = p, r
        if m == 'mDice':
            out['Dice'], out['Acc'] = 2 * ti / (tp + tl), ti / tl
    out = OrderedDict((k, v.numpy()) for k, v in out.items())
    if nan_to_num is not None:
        out = OrderedDict((k, np.nan_to_num(v, nan=nan_to_num)) for k, v in out.items())
    return out


def compute_metrics(results, metrics=('mIoU',), nan_to_num=None, beta=1):
    """iou_metric.py:120-146: float32 sums over images, nanmean*100 rounded to 2 dp."""
    cols = tuple(zip(*results))
    assert len(cols) == 4
    tot = [sum(c) for c in cols]
    ret = total_area_to_metrics(*tot, metrics, nan_to_num, beta)
    summary = {}
    for k, v in ret.items():
        val = np.round(np.nanmean(v) * 100, 2)
        summary[k if k == 'aAcc' else 'm' + k] = val
    return summary
