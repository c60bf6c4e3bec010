#!/bin/bash
# Quick GPU visit: all GPU tests + bench with per-op table.
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/pytest.log
( timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench_ops.txt )
tail -25 $O/pytest.log | cut -c1-300; cut -c1-400 $O/bench.json
