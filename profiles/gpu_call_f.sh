#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_g2_scatter -c 4 -o gpurun_out/f_scatter python profiles/coder_sweep.py 29 0.001,0.003 > gpurun_out/f_ncu_scatter.log 2>&1
echo "ncu scatter rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass_mma -c 2 -o gpurun_out/f_mma profiles/_bin/microbench > gpurun_out/f_ncu_mma.log 2>&1
echo "ncu mma rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_g2_scatter|k_g2_lengths" -c 6 -o gpurun_out/f_plane3 python profiles/profile_target.py 3 > gpurun_out/f_ncu_plane3.log 2>&1
echo "ncu plane3 rc=$?"; tail -2 gpurun_out/f_ncu_plane3.log
ls -la gpurun_out/*.ncu-rep
