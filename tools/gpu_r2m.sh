#!/bin/bash
mkdir -p gpurun_out/r2m
cd /root/repo
timeout 600 python tools/time_led_blocks.py 2>&1 | grep -v Warn | tee gpurun_out/r2m/led_blocks.txt
