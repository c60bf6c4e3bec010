"""CPU-side checks of the training host logic: flat parameter arena, the DDP gradient all-reduce over
gloo (world_size 2), PolyLR against mmengine's closed form, and that training ops refuse CPU tensors."""
import os
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lednet_b200 as L
from lednet_b200 import train_ops as T


def test_train_ops_refuse_cpu_tensors():
    with pytest.raises(L.LedB200Error):
        T.conv2d(torch.zeros(1, 4, 4, 3), torch.zeros(8, 3, 3, 3))
    with pytest.raises(L.LedB200Error):
        T.resize(torch.zeros(1, 4, 4, 3), (8, 8))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet'), dict(type='LEDHead', in_channels=128, channels=64,
                                                       num_classes=2, dropout_ratio=0.)).train()
    with pytest.raises(L.LedB200Error):
        m.loss(torch.zeros(1, 3, 64, 64), [dict(gt_sem_seg=dict(data=torch.zeros(1, 64, 64, dtype=torch.int64)))])


def test_flat_arena_views_and_zero_grad():
    lin = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    before = [p.detach().clone() for p in lin.parameters()]
    opt = L.FlatSGD(lin.parameters(), lr=0.1, world_size=1)
    assert opt.flat.numel() == sum(p.numel() for p in lin.parameters())
    for p, b in zip(lin.parameters(), before):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() >= opt.flat.data_ptr() and p.grad.data_ptr() >= opt.flat_grad.data_ptr()
    lin(torch.randn(2, 3, 8, 8)).sum().backward()          # autograd accumulates INTO the flat buffer
    assert float(opt.flat_grad.abs().sum()) > 0
    opt.zero_grad()
    assert float(opt.flat_grad.abs().sum()) == 0
    with pytest.raises(L.LedB200Error):                     # the update kernel is CUDA only
        opt.step()


def test_poly_lr_closed_form():
    class O:
        lr = 0.01
    s = L.PolyLR(O(), power=0.9, eta_min=1e-4, begin=0, end=80000)
    assert abs(s.lr_at(0) - 0.01) < 1e-12
    assert abs(s.lr_at(80000) - 1e-4) < 1e-12
    assert abs(s.lr_at(40000) - ((0.01 - 1e-4) * 0.5 ** 0.9 + 1e-4)) < 1e-12
    s.step()
    assert abs(s.opt.lr - s.lr_at(1)) < 1e-15


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Conv2d(2, 3, 3)
    opt = L.FlatSGD(net.parameters(), lr=0.1)
    assert opt.world_size == world
    opt.zero_grad()
    x = torch.full((1, 2, 5, 5), float(rank + 1))
    net(x).sum().backward()
    local = opt.flat_grad.clone()
    opt.all_reduce_grads()
    q.put((rank, local, opt.flat_grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    total = got[0][1] + got[1][1]
    assert torch.allclose(got[0][2], total) and torch.allclose(got[1][2], total)


def _eval_worker(rank, world, port, q):
    """the eval path's only collective: each rank holds the int64 (K+1) x K matrix of its own image shard"""
    import numpy as np
    import oracle
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    K = 5
    g = np.random.default_rng(100 + rank)
    pred = g.integers(0, K, (2, 24, 32))
    lab = g.integers(0, K, (2, 24, 32))
    lab[g.random(lab.shape) < 0.1] = 255
    m = L.IoUMetric(iou_metrics=['mIoU', 'mFscore'])
    m.dataset_meta = dict(classes=[str(i) for i in range(K)])
    cm = np.zeros((K + 1, K), dtype=np.int64)
    cm[:K] = oracle.confusion_matrix(pred, lab, K, 255)[:K]     # host stand-in for the CUDA histogram of this shard
    m._cm = torch.from_numpy(cm)
    local = m.total_confusion(reduce_ranks=False).clone()
    metrics = m.compute_metrics()                                # all-reduces over the two ranks
    q.put((rank, local, m.total_confusion().clone(), metrics, pred, lab))
    dist.barrier()
    dist.destroy_process_group()


def test_confusion_matrix_allreduce_gloo_world2():
    import numpy as np
    import oracle
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    total = got[0][1] + got[1][1]
    assert torch.equal(got[0][2], total) and torch.equal(got[1][2], total)
    assert got[0][3] == got[1][3]
    # and the metrics equal the reference's float32-histogram pipeline over all four images
    res = [oracle.intersect_and_union(torch.from_numpy(g[4][i]), torch.from_numpy(g[5][i]), 5, 255)
           for g in got for i in range(2)]
    ref = oracle.compute_metrics(res, ['mIoU', 'mFscore'])
    for k, v in ref.items():
        assert abs(float(got[0][3][k]) - float(v)) < 1e-9, k


def _stats_worker(rank, world, port, q):
    """SyncBN's statistics reducer off the NVLink path (gloo group): it must fall back to dist.all_reduce and sum in place"""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    red = T._PeerReduce.get(dist.group.WORLD, torch.device('cpu'))
    v = torch.arange(9, dtype=torch.float64) * (rank + 1) + 0.5
    red.reduce(v[:7])                                   # a slice, as the (sum, sumsq, count) message of a layer is
    q.put((rank, red.ok, v.tolist()))               # plain lists: no shared-memory handle to outlive this process
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_statistics_reducer_falls_back_to_the_group_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    ps = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    base = torch.arange(9, dtype=torch.float64)
    want_head = (base * 1 + 0.5) + (base * 2 + 0.5)
    for rank, ok, v in got:
        v = torch.tensor(v, dtype=torch.float64)
        assert ok is False                              # peer memory needs NCCL over GPUs of one box
        assert torch.equal(v[:7], want_head[:7])
        assert torch.equal(v[7:], (base * (rank + 1) + 0.5)[7:])   # untouched tail
