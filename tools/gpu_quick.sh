#!/bin/bash
# Quick GPU visit: conv/parity tests + bench with per-op table.
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -15 ) > $O/pytest.log
( timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench_ops.txt )
tail -5 $O/pytest.log; cat $O/bench.json
