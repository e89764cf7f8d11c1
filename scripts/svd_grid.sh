for ratio in 1e-2 1 1e9; do for inner in 1 2; do for inner0 in 2 3; do
echo "ratio=$ratio inner=$inner inner0=$inner0"
QM_EIG_OLD=1 QM_SVD_CROSS_RATIO=$ratio QM_SVD_INNER=$inner QM_SVD_INNER0=$inner0 python scripts/svd_probe.py 512 2048 1024 1024 2>&1 | grep svd
done; done; done
