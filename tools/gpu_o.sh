#!/bin/bash
# round 2, call O: timing of the line-search pieces (debug bits 64 + 16) on one window with failing searches
cd /root/repo
mkdir -p gpurun_out
VRF_BA_DEBUG=80 timeout 300 python /root/repo/tools/scratch/ls_timing.py > gpurun_out/o_ls.log 2>&1
grep -c "ls trial" gpurun_out/o_ls.log
grep "ls step\|ls trial\|evaluate mode=2 tid=0 " gpurun_out/o_ls.log | head -40
