#!/bin/bash
# round 2, call l: LEDNet(variant='led')
mkdir -p gpurun_out/r2l
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_led_variant.py -x -q -m gpu -s > gpurun_out/r2l/led.log 2>&1; echo "led rc=$?"
grep -E "c5|passed|failed|Error|error|assert" gpurun_out/r2l/led.log | tail -30
timeout 600 python bench.py --variant led --steps 5 --warmup 3 > gpurun_out/r2l/bench_led.json 2> gpurun_out/r2l/bench_led.err; echo "bench led rc=$?"
tail -3 gpurun_out/r2l/bench_led.err; cut -c1-400 gpurun_out/r2l/bench_led.json
timeout 600 python bench.py --variant led --dtype fp32 --steps 5 --warmup 3 2>/dev/null | cut -c1-300 | tee gpurun_out/r2l/bench_led_fp32.json
