#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
( time python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --workload c5 --batch 256 --lanes 8 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python scripts/profile_stages.py --n 12 --chi 64 --layers 10 --sweeps 20 > gpurun_out/stages_c5.log 2>&1
python scripts/bench_kernels.py gpurun_out/bench_kernels.json > gpurun_out/bench_kernels.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 8000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
gzip -f gpurun_out/launches_bench.csv
cut -c1-300 gpurun_out/bench_final.json; cat gpurun_out/bench_ref.json | cut -c1-600; tail -3 gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_c5.json; cut -c1-300 gpurun_out/bench_c2.json; cat gpurun_out/stages_c5.log | tail -14
