#!/bin/bash
# A/B of the split-K rule of qm_zgemm (QM_GEMM_SPLITK_MINK=512: round-2 rule) at mid-size shapes vs cuBLAS, then the
# whole GPU suite and the config-2 line with the new default.
mkdir -p gpurun_out
{
echo "== QM_GEMM_SPLITK_MINK=512 (old rule)"; QM_GEMM_SPLITK_MINK=512 python scripts/bench_kernels.py gpurun_out/kernels_gemm_old.json --no-svd 2>&1 | tail -16
echo "== default (k >= 256, slices >= 64)"; python scripts/bench_kernels.py gpurun_out/kernels_gemm_new.json --no-svd 2>&1 | tail -16
} > gpurun_out/gemm_splitk_ab.log 2>&1
cat gpurun_out/gemm_splitk_ab.log | cut -c1-220
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2_pytest54.log; cat gpurun_out/r2_pytest54.log
python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench54_c2.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench54_c2.json')); print('c2 ms', d['ms_per_step'], d['fidelity_mean'], d['roofline'].get('classes',{}).get('zgemm'))"
