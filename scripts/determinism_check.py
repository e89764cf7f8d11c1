"""Run the same prepare twice (and optionally compare with a saved run): detects races in the launch
schedule (programmatic dependent launch, two-stream Jacobi schedule).  GPU box.
usage: determinism_check.py n chi layers sweeps out.npy [ref.npy]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qmprs_b200 import host
from qmprs_b200.kernels import get_kernels
n, chi, L, S = [int(x) for x in sys.argv[1:5]]
K = get_kernels("cuda:0")
rng = np.random.default_rng(3)
v = rng.random(2 ** n) + 1j * rng.random(2 ** n); v /= np.linalg.norm(v)
runs = []
for rep in range(3):
    r = host.prepare(K, K.from_host(v), n, chi, L, S)
    runs.append((np.asarray(r["gates"]).copy(), r["fidelity"]))
g0, f0 = runs[0]
for g, f in runs[1:]:
    print("repeat: max|dgates| %.3e  dfid %.3e" % (np.abs(g - g0).max(), abs(f - f0)))
np.save(sys.argv[5], g0)
if len(sys.argv) > 6:
    ref = np.load(sys.argv[6])
    print("vs ref: max|dgates| %.3e" % np.abs(ref - g0).max(), "fidelity", f0)
