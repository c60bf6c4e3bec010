#!/bin/bash
mkdir -p gpurun_out/r2c
cd /root/repo
timeout 300 python tools/diag_ladder5.py 2>&1 | grep -v Warning | tee -a gpurun_out/r2c/diag5.txt
LEDB200_POISON=127 timeout 300 python tools/diag_ladder5.py 2>&1 | grep -v Warning | tee -a gpurun_out/r2c/diag5.txt
LEDB200_POISON=255 timeout 300 python tools/diag_ladder5.py 2>&1 | grep -v Warning | tee -a gpurun_out/r2c/diag5.txt
