#!/usr/bin/env python
"""Bare pinned host -> device bandwidth with every rank uploading at once: the ceiling of the
end-to-end figure at N GPUs.  Each rank copies the bench's per-step upload (16 x 1024 x 2048 uint8 BGR images + labels =
134 MB) from pinned host memory on its own stream, back to back for ~1.5 s; reports GB/s per rank, the aggregate, the
step rate that bandwidth alone would allow, the CPU affinity / NUMA node of every rank and (optionally) the same with
the process bound to the GPU's NUMA-local cores.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/probe_h2d.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def numa_of_gpu(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        p = f'/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node'
        return int(open(p).read()) if os.path.exists(p) else None
    except Exception:
        return None


def run(nbytes, seconds, dev):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    devb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        devb.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            devb.copy_(host, non_blocking=True)
        n += 8
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return nbytes * n / (ms * 1e-3) / 1e9


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    nbytes = 16 * 1024 * 2048 * 4                       # 3 image bytes + 1 label byte per pixel
    aff = sorted(os.sched_getaffinity(0))
    gbs = run(nbytes, 1.5, dev)
    numa = numa_of_gpu(local)
    # second pass: bound to the GPU's NUMA-local cores when the node is known and differs from "everything"
    bound = None
    if numa is not None and numa >= 0:
        try:
            cpus = open(f'/sys/devices/system/node/node{numa}/cpulist').read().strip()
            ids = []
            for part in cpus.split(','):
                a, _, b = part.partition('-')
                ids += list(range(int(a), int(b or a) + 1))
            ids = [c for c in ids if c in aff]
            if ids and len(ids) < len(aff):
                os.sched_setaffinity(0, ids)
                bound = run(nbytes, 1.5, dev)
        except Exception:
            bound = None
    rows = [None] * world
    dist.all_gather_object(rows, dict(rank=rank, gbs=gbs, gbs_numa_bound=bound, numa_node=numa, n_affinity=len(aff),
                                      affinity=f'{aff[0]}-{aff[-1]}'))
    if rank == 0:
        total = sum(r['gbs'] for r in rows)
        out = dict(probe='pinned H2D, all ranks at once', n_gpus=world, bytes_per_copy=nbytes,
                   aggregate_gbs=total, per_rank_gbs=[round(r['gbs'], 2) for r in rows],
                   per_rank_gbs_numa_bound=[r['gbs_numa_bound'] and round(r['gbs_numa_bound'], 2) for r in rows],
                   upload_only_steps_per_s_per_rank=min(r['gbs'] for r in rows) * 1e9 / nbytes,
                   upload_only_img_per_s=16 * world * min(r['gbs'] for r in rows) * 1e9 / nbytes,
                   numa_nodes=[r['numa_node'] for r in rows], affinity=[r['affinity'] for r in rows],
                   host_cpus=os.cpu_count())
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
