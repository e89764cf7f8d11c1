"""Which SVDs the headline path runs and how many Jacobi sweeps each takes.
usage: python scripts/svd_census.py [n_qubits=20] [chi=512] [layers=15]"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qmprs_b200 import host
from qmprs_b200.kernels import get_kernels

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 20
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 512
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 15
K = get_kernels()
rng = np.random.default_rng(0)
psi = rng.standard_normal(2 ** nq) + 1j * rng.standard_normal(2 ** nq)
psi /= np.linalg.norm(psi)
K.svd_log = []
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = host.prepare(K, psi, nq, chi, layers, 0)
e1.record(); torch.cuda.synchronize()
print(f"{nq} q chi={chi} {layers} layers, no sweeps: {e0.elapsed_time(e1):.1f} ms, fidelity {res['fidelity']:.6f}")
agg = collections.OrderedDict()
for m, n, sw, bm in K.svd_log:
    a = agg.setdefault((m, n), [0, 0, 0, 99])
    a[0] += 1; a[1] += sw; a[2] = max(a[2], sw); a[3] = min(a[3], sw)
print("shape: count, total sweeps, min..max sweeps, rounds per sweep (pairs of 16-row blocks)")
tot = 0
for (m, n), (c, s, mx, mn) in sorted(agg.items(), key=lambda kv: -kv[1][1] * min(kv[0]) * max(kv[0])):
    nb = (min(m, n) + 15) // 16
    print(f"  {m:5d} x {n:5d}: {c:4d} SVDs, {s:5d} sweeps ({mn}..{mx}), {max(nb - 1, 1)} rounds/sweep")
    tot += s
print("multi-CTA SVDs:", len(K.svd_log), "sweeps:", tot)
seq = [(m, n, sw) for m, n, sw, _ in K.svd_log]
print("first 40 in order:", seq[:40])
