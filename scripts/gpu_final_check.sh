#!/bin/bash
# Round-end check on one GPU box: the whole -m gpu suite, the default bench line, the config-5 line, smoke().
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2_pytest50.log; echo "pytest rc=$?"; cat gpurun_out/r2_pytest50.log
python bench.py > gpurun_out/r2_bench50.json 2> gpurun_out/r2_bench50.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench50.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['fidelity_mean'], d['cpu_baseline']['value'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['clocks'])
b=d['batch_c5']; print(b['value'], b['e2e']['value'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
