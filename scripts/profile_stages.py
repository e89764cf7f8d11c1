"""Stage timing + per-kernel-class CUDA-event profile of one prepare_state (GPU box)."""
import argparse, json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200 import host
from qmprs_b200.kernels import get_kernels

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16); ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--layers", type=int, default=3); ap.add_argument("--sweeps", type=int, default=2)
ap.add_argument("--prof", type=int, default=1)
a = ap.parse_args()
K = get_kernels("cuda:0")
rng = np.random.default_rng(0)
v = rng.random(2 ** a.n) + 1j * rng.random(2 ** a.n); v /= np.linalg.norm(v)
psi = K.from_host(v)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
# warm
host.prepare(K, K.from_host(v[: 2 ** 8] / np.linalg.norm(v[:2**8])), 8, 16, 1, 1)
if a.prof: K.prof_begin()
t0 = T(); rec = {}
A = host.build_mps(K, psi, a.n, a.chi, rec); t1 = T()
print("build_mps", round(t1 - t0, 4), "bonds", host.bond_dims(A), "svd_sweeps", K.svd_sweeps)
B = host.copy_mps(K, A); host.normalize_site0(K, B)
for l in range(a.layers):
    s0 = K.svd_sweeps; t2 = T(); g, k = host.chi2_layer(K, B); t3 = T(); host.apply_inverse_layer(K, B, g, k); t4 = T(); f = host.zero_overlap(K, B); t5 = T()
    print(f"layer {l}: chi2 {t3-t2:.4f}  apply_inverse {t4-t3:.4f}  overlap {t5-t4:.4f}  f={abs(f):.6f} svd_sweeps {K.svd_sweeps-s0} bonds max {max(host.bond_dims(B))}")
gates = torch.cat([g] * a.layers); kinds = [k] * a.layers
t6 = T(); target = host.to_dense(K, A); t7 = T()
host.optimize_layers(K, target, gates, kinds, a.n, a.sweeps); t8 = T()
print(f"to_dense {t7-t6:.4f}  sweeps {a.sweeps}: {t8-t7:.4f}  per sweep {(t8-t7)/max(a.sweeps,1):.4f}")
if a.prof:
    p = K.prof_end()
    for kname, (ms, cnt, work) in sorted(p.items(), key=lambda kv: -kv[1][0]):
        print(f"  {kname:12s} {ms:10.2f} ms  {cnt:8d} launches  {1e3*ms/max(cnt,1):8.2f} us/launch  work {work:.3e}  rate {work/max(ms,1e-9)/1e9:.1f} G/s")
