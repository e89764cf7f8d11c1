#!/bin/bash
# One GPU-box call: new parity tests, A/B of the k_round Gram phase (QM_SVD_GRAM2), config-5 bench line.
mkdir -p gpurun_out
python -m pytest tests/test_api_gpu.py tests/test_graphs_gpu.py tests/test_headline_gpu.py -q -m gpu -k "iterative or batch or graph or config2" -x 2>&1 | tail -15 > gpurun_out/r2_pytest_sched.log
cat gpurun_out/r2_pytest_sched.log
for g in 0 1; do
  echo "== QM_SVD_GRAM2=$g"
  QM_SVD_GRAM2=$g python scripts/svd_probe.py 1024 1024 512 2048 2048 512 256 1024 256 256 128 512 2>&1 | tail -8
  QM_SVD_GRAM2=$g QM_ROUND_DEBUG=1 python scripts/svd_probe.py 1024 1024 512 2048 2>&1 | grep -m2 "k_round"
done > gpurun_out/svd_gram2_ab.log 2>&1
cat gpurun_out/svd_gram2_ab.log
python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench49_c5.json 2> gpurun_out/r2_bench49_c5.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench49_c5.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
