#!/bin/bash
# Scan of QM_SVD_EARLY (a Jacobi sweep that STARTS below this |cos| is the last one): sweeps, time and accuracy on the
# gate-split shapes, then the parity suites and the headline step under two looser values.
mkdir -p gpurun_out
{
for e in 1e-9 1e-7 1e-6 1e-5 1e-4 1e-3; do
  echo "== QM_SVD_EARLY=$e"
  QM_SVD_EARLY=$e QM_PROBE_BACKMULT=1 python scripts/svd_probe.py 1024 1024 512 2048 256 1024 128 512 2>&1 | tail -4
done
for e in 1e-6 1e-4; do
  echo "== parity suites with QM_SVD_EARLY=$e"
  QM_SVD_EARLY=$e python -m pytest tests/test_pipeline_gpu.py tests/test_headline_gpu.py tests/test_api_gpu.py tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -12
  QM_SVD_EARLY=$e python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('headline ms', d['ms_per_step'], 'fidelity', d['fidelity_mean'], 'launches', d['gpu_launches'])"
done
} > gpurun_out/svd_early_scan.log 2>&1
cat gpurun_out/svd_early_scan.log
