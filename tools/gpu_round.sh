#!/bin/bash
# One GPU-box visit: parity tests, bench (with per-op table), ncu launch list, ncu full capture of the top kernels.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > $O/smoke.log
( timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench_ops.txt )
( timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err )
# launch list (cold-cache, serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $O/launches.csv python tools/prof_forward.py --batch 4 --iters 2 > $O/launches.log 2>&1
# full capture of the dominant kernel family (conv_tc), tail and stem
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 12 \
   -o $O/conv_tc python tools/prof_forward.py --batch 4 --iters 2 > $O/ncu_conv_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tail|stem_tc|upsample|confusion' -s 6 -c 8 \
   -o $O/others python tools/prof_forward.py --batch 4 --iters 2 > $O/ncu_others.log 2>&1
ls -la $O
cat $O/pytest_gpu.log | tail -8; cat $O/smoke.log; cat $O/bench.json
