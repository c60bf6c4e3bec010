#!/usr/bin/env python
"""Training-step benchmark (BASELINE config 4): 1024x1024 crop, batch 12/GPU, K=19, forward + backward +
OHEM CE + SGD, DDP gradient all-reduce when launched under torchrun.  Prints one JSON line (rank 0).

    python tools/bench_train.py [--batch 12] [--size 1024] [--steps 5] [--warmup 2]
"""
import argparse
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=12)
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--classes', type=int, default=19)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--syncbn', action='store_true', help="norm_cfg=dict(type='SyncBN') as in the reference config (N > 1)")
    args = ap.parse_args()
    import torch
    import lednet_b200 as L
    from lednet_b200 import synth
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    K, N, S = args.classes, args.batch, args.size
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        norm = dict(type='SyncBN', requires_grad=True) if args.syncbn else dict(type='BN', requires_grad=True)
        m = L.EncoderDecoder(dict(type='LEDNet', norm_cfg=norm),
                             dict(type='LEDHead', in_channels=128, channels=64, num_classes=K, dropout_ratio=0., norm_cfg=norm),
                             data_preprocessor=None, compute_dtype='fp32')
    m.load_state_dict(synth.make_state_dict(m.state_dict(), seed=2))
    m.to(dev).train()
    opt = L.FlatSGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    sched = L.PolyLR(opt, power=0.9, eta_min=1e-4, end=80000)
    img = synth.make_images_u8(N, S, S, seed=100 + rank).to(dev)
    mean = torch.tensor(L.engine.MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(L.engine.STD, device=dev).view(1, 3, 1, 1)
    x = ((img[:, [2, 1, 0]].float() - mean) / std).contiguous()
    lab = synth.make_labels(N, S, S, K, seed=200 + rank).to(dev)
    samples = [dict(gt_sem_seg=dict(data=lab[i:i + 1])) for i in range(N)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phases = [0.0, 0.0, 0.0]

    def step(timed):
        ev[0].record()
        total, log = m.parse_losses(m.loss(x, samples))
        ev[1].record()
        opt.zero_grad()
        total.backward()
        ev[2].record()
        opt.step()
        sched.step()
        ev[3].record()
        if timed:
            torch.cuda.synchronize()
            for i in range(3):
                phases[i] += ev[i].elapsed_time(ev[i + 1])
        return log

    for _ in range(args.warmup):
        log = step(False)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        log = step(True)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / args.steps
    if rank == 0:
        print(json.dumps(dict(metric='LED-Net train img/s @1024x1024', value=world * N / (ms_step * 1e-3),
                              unit='img/s', n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                              dtype=L.train_ops.compute_mode(), norm='SyncBN' if args.syncbn else 'BN', data='synthetic', loss=float(log['loss'].detach()),
                              phases_ms=dict(forward_loss=phases[0] / args.steps, backward=phases[1] / args.steps,
                                             allreduce_sgd=phases[2] / args.steps),
                              peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
                              config=dict(workload='LED-Net(R0)+LEDHead train step: fwd + OHEM CE x2 + bwd + SGD',
                                          batch_per_gpu=N, height=S, width=S, num_classes=K))))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
