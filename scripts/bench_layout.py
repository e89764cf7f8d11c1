"""HBM bandwidth of the TT-SVD layout kernels (qm_transpose) at the 24-qubit split shapes, and the
time of the first (skinny) TT-SVD splits.  Algorithmic bytes: 32 per element (read + write)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200.kernels import get_kernels

K = get_kernels("cuda:0")
peak = 6443.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows_out = []
for rows, cols in [(1 << 23, 2), (1 << 22, 4), (1 << 20, 16), (1 << 19, 32), (1 << 16, 256), (1 << 13, 2048), (4096, 4096),
                   (2, 1 << 23), (32, 1 << 19)]:
    a = torch.randn(rows, cols, dtype=torch.complex128, device="cuda:0")
    for _ in range(3):
        K.transpose(a)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    ts = []
    for i in range(5):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.transpose(a); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gbs = 32.0 * rows * cols / (ms * 1e-3) / 1e9
    rows_out.append({"rows": rows, "cols": cols, "ms": ms, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak})
    print(f"transpose {rows:8d} x {cols:8d}: {ms:8.3f} ms  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f}% of {peak:.0f}")
# first TT-SVD splits at 24 qubits (SVD of 2^i x 2r, r = 1,2,4,...)
rng = np.random.default_rng(0)
v = torch.from_numpy(rng.random(1 << 24) + 1j * rng.random(1 << 24)).to("cuda:0")
for i, r in [(23, 1), (22, 2), (21, 4), (20, 8), (19, 16)]:
    M = v.reshape(1 << i, 2 * r)
    K.svd(M)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.svd(M); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"svd split i={i} ({1<<i} x {2*r}): {ms:8.3f} ms  ({256e6/ (ms*1e-3)/1e9:.0f} GB/s counted as one pass over the 256 MB vector)")
    rows_out.append({"svd_split": i, "cols": 2 * r, "ms": ms})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows_out, open("gpurun_out/layout_bw.json", "w"), indent=1)
