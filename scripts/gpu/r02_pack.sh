#!/bin/bash
# k_pack after the full-slot fast path: variants, then one ncu --set full capture and the GPU parity tests
mkdir -p gpurun_out
timeout 300 python scripts/gpu/r02_pack.py 2>&1 | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pack -s 1 -c 1 -o gpurun_out/r02_pack python scripts/gpu/r02_pack.py default > gpurun_out/r02_pack_ncu.log 2>&1
ncu -i gpurun_out/r02_pack.ncu-rep --page details > gpurun_out/r02_pack_details.txt 2>&1
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|DRAM Throughput|Executed Ipc Active|Issue Slots Busy|Waves Per SM|Block Limit" gpurun_out/r02_pack_details.txt | head -20
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
