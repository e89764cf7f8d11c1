#!/bin/bash
# round-1 re-entry: validate restored tree, refresh bench + ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python scripts/profile_stages.py --n 20 --chi 512 --layers 2 --sweeps 2 > gpurun_out/stages_c3.log 2>&1
python scripts/svd_probe.py 1024 1024 512 2048 256 256 128 512 64 64 > gpurun_out/svd_probe.log 2>&1
for spec in "k_eig 40 eig_early" "k_eig 1500 eig_late" ; do
  set -- $spec
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 python scripts/svd_probe.py 1024 1024 > gpurun_out/ncu_$3.log 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_env_fused -s 100 -c 1 -f -o gpurun_out/env_fused python scripts/profile_stages.py --n 20 --chi 512 --layers 1 --sweeps 1 --prof 0 > gpurun_out/ncu_env.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_zgemm_tma -s 2 -c 1 -f -o gpurun_out/zgemm_tma python scripts/bench_kernels.py > gpurun_out/ncu_zgemm.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_default.json | cut -c1-600
