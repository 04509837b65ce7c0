#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv1x1_chain_nhwc' -s 2 -c 2 -f -o gpurun_out/r1s_prof python tools/profile_chain.py > gpurun_out/r1s_ncu.log 2>&1
tail -3 gpurun_out/r1s_ncu.log
