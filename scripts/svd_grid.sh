# schedule-knob scan for the Jacobi SVD (GPU box): cross-only threshold of the inner eigen-solve
for ratio in 1e-2 1 10 100 1e4; do
echo "ratio=$ratio"
QM_SVD_CROSS_RATIO=$ratio python scripts/svd_probe.py 256 1024 512 2048 1024 1024 2>&1 | grep svd
done
