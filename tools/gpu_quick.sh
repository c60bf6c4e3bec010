#!/bin/bash
# Quick GPU visit: conv/parity/block tests + bench with per-op table.
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_parity.py tests/test_gpu_blocks.py -q 2>&1 | tail -40 ) > $O/pytest.log
( timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench_ops.txt )
tail -25 $O/pytest.log | cut -c1-300; cut -c1-400 $O/bench.json
