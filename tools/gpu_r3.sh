#!/bin/bash
# Round-3 evidence on one GPU: everything gpu_final.sh records for the inference headline, plus the training step
# (bench_train in its three arithmetic modes, launch list, full ncu captures of the tensor-core training kernels).
TAG=${1:-r3z}
O=gpurun_out/$TAG
mkdir -p $O
cd /root/repo
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ); echo "bench rc=$?"
( timeout 400 python bench.py --steps 20 --warmup 5 --profile-ops --no-extras > $O/bench_ops.json 2> $O/bench_ops.txt )
( timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err )
# training step: three-pass tensor cores (default), single pass, CUDA cores
( timeout 300 python tools/bench_train.py --steps 5 --warmup 2 > $O/train_tf32x3.json 2> $O/train.err )
( LEDB200_TRAIN_TC=fast timeout 300 python tools/bench_train.py --steps 5 --warmup 2 > $O/train_tf32.json 2>> $O/train.err )
( LEDB200_TRAIN_TC=0 timeout 300 python tools/bench_train.py --steps 3 --warmup 1 > $O/train_f32.json 2>> $O/train.err )
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv \
   python tools/bench_train.py --steps 1 --warmup 1 > $O/train_launches.log 2>&1
python tools/summarize_launches.py $O/train_launches.csv 2 > $O/train_launches.txt 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none -k regex:wgrad_tc_kernel -s 8 -c 3 \
   -o $O/wgrad_tc python tools/bench_train.py --steps 1 --warmup 0 > $O/ncu_wgrad.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 10 -c 4 \
   -o $O/conv_tc_tf32 python tools/bench_train.py --steps 1 --warmup 0 > $O/ncu_conv_tf32.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'ohem_up|chan_reduce4|bn_bwd_apply4|resize_bwd_row|wgrad_small' -c 12 \
   -o $O/train_others python tools/bench_train.py --steps 1 --warmup 0 > $O/ncu_train_others.log 2>&1
# inference: launch list + traffic of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/traffic.csv python tools/prof_forward.py --batch 16 --iters 1 > $O/traffic.log 2>&1
ls -la $O
cut -c1-600 $O/bench.json; echo
for f in train_tf32x3 train_tf32 train_f32; do cut -c1-330 $O/$f.json; echo; done
head -30 $O/train_launches.txt
