#!/bin/bash
# A/B of k_round Gram-phase variants on ONE box, SM clocks sampled during each run.
# ctl = library built from the previous commit's svd.cu (qmprs_b200/libqmprs_b200_ctl.so), new = in-tree library.
# Control build (here, before the gpurun call):
#   git show HEAD:qmprs_b200/csrc/svd.cu > /tmp/ctl/svd.cu
#   nvcc <FLAGS of qmprs_b200/build.py> <the other csrc/*.cu> /tmp/ctl/svd.cu -o qmprs_b200/libqmprs_b200_ctl.so
# The default probe accumulates V in the identity extension (update over 2 x len columns); QM_PROBE_BACKMULT=1 is the
# mode the pipeline uses (V by back-multiplication): 1024 x 1024 takes 56 ms in the former, 44-46 ms in the latter.
mkdir -p gpurun_out
run() {  # name lib gram2 [extra env]
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader,nounits -lms 100 > /tmp/smi_$1.txt &
  local pid=$!
  echo "== $1"
  env QM_B200_LIB=$2 QM_SVD_GRAM2=$3 $4 python scripts/svd_probe.py 1024 1024 512 2048 256 1024 2>&1 | grep -v "^$" | tail -6
  kill $pid
  sort -t, -k1 -n /tmp/smi_$1.txt | awk -F, '{c[NR]=$1; p[NR]=$2} END {print "   sm clock MHz min/median/max:", c[1], c[int((NR+1)/2)], c[NR], " samples", NR}'
  sort -t, -k2 -n /tmp/smi_$1.txt | awk -F, '{p[NR]=$2} END {print "   power W median/max:", p[int((NR+1)/2)], p[NR]}'
  cut -d, -f3 /tmp/smi_$1.txt | sort | uniq -c | head -4
}
CTL=$PWD/qmprs_b200/libqmprs_b200_ctl.so
NEW=$PWD/qmprs_b200/libqmprs_b200.so
{
run ctl $CTL 0
run new_gram0 $NEW 0
run new_gram1 $NEW 1
run new_gram2 $NEW 2
run ctl_again $CTL 0
run ctl_dbg $CTL 0 QM_ROUND_DEBUG=1
run new_gram2_dbg $NEW 2 QM_ROUND_DEBUG=1
} > gpurun_out/svd_gram2_ab2.log 2>&1
cat gpurun_out/svd_gram2_ab2.log
python -m pytest tests/test_headline_gpu.py -q -m gpu -k "config2" -x 2>&1 | tail -15 > gpurun_out/r2_pytest_c2.log
cat gpurun_out/r2_pytest_c2.log
