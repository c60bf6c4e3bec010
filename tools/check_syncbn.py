#!/usr/bin/env python
"""SyncBN exactness check, launched with torchrun --nproc-per-node 2 (both ranks may share ONE GPU: gloo backend).

Op level: T.bn_act with `sync` on half batches of two ranks == torch BatchNorm2d (float64, CPU) on the full batch:
output, d/dx, running statistics; per-rank dgamma/dbeta sum to the full-batch ones (torch.nn.SyncBatchNorm).
Model level: one FlatSGD training step of LEDNet+LEDHead built with norm_cfg SyncBN on 2 ranks; the SyncBN layers'
running statistics are identical on both ranks afterwards, the plain-BN layers' (first block of each layer, DAPPM: the
reference builds those without norm_cfg) are not.
"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    import lednet_b200 as L
    from lednet_b200 import synth, train_ops as T
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(rank % ngpu)
    dev = torch.device('cuda', rank % ngpu)
    dist.init_process_group('gloo' if ngpu < world else 'nccl')
    ok = True
    if dist.get_backend() == 'nccl':
        # the peer-memory all-reduce of the statistics (ledb200_peer_allreduce_f64) against NCCL's, several rounds so both
        # parities of its double-buffered slots and flags are exercised; identical bits on every rank
        pr = T._PeerReduce.get(dist.group.WORLD, dev)
        for it in range(5):
            v = (torch.arange(517, dtype=torch.float64, device=dev) * (rank + 1.25) + it) / 7.0
            ref = v.clone()
            dist.all_reduce(ref)
            pr.reduce(v)
            same = torch.allclose(v, ref, rtol=1e-15, atol=0)
            gathered = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(gathered, v)
            ident = all(torch.equal(gathered[0], t) for t in gathered)
            ok = ok and pr.ok and same and ident
        if rank == 0:
            print(f'peer all-reduce: in use {pr.ok}, matches NCCL {same}, identical on all ranks {ident} -> {"ok" if ok else "FAIL"}')
    g = torch.Generator().manual_seed(5)
    for (c, hw, relu, res) in [(32, (12, 20), True, True), (19, (9, 7), True, False), (64, (4, 4), False, False)]:
        n = 2 * world
        bn = torch.nn.BatchNorm2d(c).double()
        with torch.no_grad():
            bn.weight.copy_(torch.rand(c, generator=g) + 0.5); bn.bias.copy_(torch.randn(c, generator=g) * 0.1)
        y = (torch.randn(n, c, *hw, generator=g) * 2 + 0.5).double().requires_grad_()
        r = torch.randn(n, c, *hw, generator=g).double() if res else None
        dout = torch.randn(n, c, *hw, generator=g).double()
        ref = bn(y)
        if res:
            ref = ref + r
        if relu:
            ref = F.relu(ref)
        ref.backward(dout)
        sl = slice(2 * rank, 2 * rank + 2)
        bnd = torch.nn.BatchNorm2d(c).to(dev)
        with torch.no_grad():
            bnd.weight.copy_(bn.weight.float()); bnd.bias.copy_(bn.bias.float())
        bnd.sync = True
        nhwc = lambda t: t.float().permute(0, 2, 3, 1).contiguous().to(dev)      # noqa: E731
        yd = nhwc(y.detach()[sl]).requires_grad_()
        out = T.bn_act(yd, bnd, res=nhwc(r[sl]) if res else None, relu=relu)
        out.backward(nhwc(dout[sl]))
        rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))   # noqa: E731
        e_out = rel(out.detach().permute(0, 3, 1, 2), ref.detach()[sl])
        e_dx = rel(yd.grad.permute(0, 3, 1, 2), y.grad[sl])
        e_rm = rel(bnd.running_mean, bn.running_mean)
        e_rv = rel(bnd.running_var, bn.running_var)
        dg = bnd.weight.grad.clone(); db = bnd.bias.grad.clone()
        dist.all_reduce(dg); dist.all_reduce(db)
        e_dg, e_db = rel(dg, bn.weight.grad), rel(db, bn.bias.grad)
        good = max(e_out, e_rm, e_rv) < 1e-5 and max(e_dx, e_dg, e_db) < 2e-4
        ok = ok and good
        if rank == 0:
            print(f'syncbn op C={c} hw={hw}: out {e_out:.1e} dx {e_dx:.1e} dgamma {e_dg:.1e} dbeta {e_db:.1e} '
                  f'running mean {e_rm:.1e} var {e_rv:.1e} -> {"ok" if good else "FAIL"}')
    # ---- model level
    K = 3
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = L.EncoderDecoder(dict(type='LEDNet', norm_cfg=dict(type='SyncBN', requires_grad=True)),
                             dict(type='LEDHead', in_channels=128, channels=64, num_classes=K, dropout_ratio=0.,
                                  norm_cfg=dict(type='SyncBN', requires_grad=True)),
                             data_preprocessor=None, compute_dtype='fp32')
    m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
    m.to(dev).train()
    opt = L.FlatSGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    img, lab = synth.make_scene(2, 128, 128, K, seed=50 + rank, coarse=(4, 4))
    x = ((img[:, [2, 1, 0]].float() - torch.tensor(L.engine.MEAN).view(1, 3, 1, 1)) / torch.tensor(L.engine.STD).view(1, 3, 1, 1)).to(dev)
    samples = [dict(gt_sem_seg=dict(data=lab[i:i + 1].to(dev))) for i in range(2)]
    total, log = m.parse_losses(m.loss(x, samples))
    opt.zero_grad(); total.backward(); opt.step()
    n_sync = n_plain = bad_sync = differing_plain = 0
    for name, mod in m.named_modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            rm = mod.running_mean.detach().clone()
            lo, hi = rm.clone(), rm.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool((lo == hi).all())
            if getattr(mod, 'sync', False):
                n_sync += 1; bad_sync += (not same)
            else:
                n_plain += 1; differing_plain += (not same)
    good = bad_sync == 0 and n_sync >= 30 and n_plain > 10
    ok = ok and good and bool(torch.isfinite(total))
    if rank == 0:
        print(f'syncbn model: loss {float(total.detach()):.4f}; {n_sync} SyncBN layers, {bad_sync} with rank-dependent running stats; '
              f'{n_plain} plain BN layers ({differing_plain} rank-dependent, as in the reference) -> {"ok" if good else "FAIL"}')
    flag = torch.tensor([0 if ok else 1], device=dev if dist.get_backend() == 'nccl' else 'cpu')
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() > 0))


if __name__ == '__main__':
    main()
