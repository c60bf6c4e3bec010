#!/bin/bash
# 8-GPU: bare pinned-H2D ceiling, then the bench (device-resident + e2e) on the final tree
mkdir -p gpurun_out/scale8
cd /root/repo
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/probe_h2d.py 2>gpurun_out/scale8/probe.err | tee gpurun_out/scale8/h2d_probe_${N}gpu.json
tail -3 gpurun_out/scale8/probe.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale8/bench.err > gpurun_out/scale8/bench_${N}gpu.json
python - <<PY
import json
d=json.loads(open('gpurun_out/scale8/bench_${N}gpu.json').read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d.get('sustained',{}).get('value'))
print('extras', {k:(v or {}).get('value') for k,v in (d.get('extra') or {}).items()})
PY
tail -3 gpurun_out/scale8/bench.err
nvidia-smi topo -m > gpurun_out/scale8/topo.txt 2>&1; lscpu | head -25 > gpurun_out/scale8/lscpu.txt; numactl -H >> gpurun_out/scale8/lscpu.txt 2>&1
