#!/bin/bash
# Round-end evidence on one GPU: bench (with extras) + per-op table + reference arm, the ncu launch list of the bench
# command, per-launch DRAM traffic of one forward at the bench workload, full ncu captures of the top kernels.
TAG=${1:-r2z}
O=gpurun_out/$TAG
mkdir -p $O
cd /root/repo
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 600 python bench.py > $O/bench.json 2> $O/bench.err ); echo "bench rc=$?"
( timeout 400 python bench.py --steps 20 --warmup 5 --profile-ops --no-extras > $O/bench_ops.json 2> $O/bench_ops.txt )
( timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err )
( timeout 300 python bench.py --variant led --steps 10 --warmup 3 > $O/bench_led.json 2> $O/bench_led.err )
# launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/launches.log 2>&1
# DRAM bytes per launch, one forward at the bench workload (batch 16)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/traffic.csv python tools/prof_forward.py --batch 16 --iters 1 > $O/traffic.log 2>&1
# full captures
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 15 -c 2 \
   -o $O/conv_tc_spa python tools/prof_forward.py --batch 16 --iters 1 > $O/ncu_spa.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder -c 2 \
   -o $O/ladder python tools/prof_forward.py --batch 16 --iters 1 > $O/ncu_ladder.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'dappm|stem_tc|upsample_add16|confusion' -c 6 \
   -o $O/others python tools/prof_forward.py --batch 16 --iters 1 > $O/ncu_others.log 2>&1
ls -la $O
cut -c1-900 $O/bench.json; echo; cut -c1-300 $O/bench_ref.json; echo; cut -c1-300 $O/bench_led.json; echo
for f in launches traffic ncu_spa ncu_ladder ncu_others; do tail -n 2 $O/$f.log; done
