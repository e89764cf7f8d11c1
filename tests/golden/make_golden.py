"""Writes tests/golden/oracle_canonical.npz from the oracle (canonical gauge).

The reference cannot be imported here (quimb / quick absent), so these
fixtures pin the ORACLE, and through it the CUDA path on the GPU box, not the real quimb
path.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qmprs_oracle as O  # noqa: E402

CASES = {"c6": (6, 64, 3, 4, 11), "c8": (8, 32, 5, 3, 12), "c10": (10, 512, 6, 2, 13)}
out = {}
for key, (n, chi, L, S, seed) in CASES.items():
    psi = O.random_state(n, seed)
    rec = {}
    res = O.prepare(psi, n, chi, L, S, gauge="canonical", record=rec)
    g = np.zeros((L, n, 16), dtype=complex)
    k = np.zeros((L, n), dtype=np.int32)
    for li, _, _, site, G in O.flatten_layers(res["layers"]):
        g[li, site, : G.size] = G.reshape(-1)
        k[li, site] = 2 if G.shape[0] == 4 else 1
    out[key + "_cfg"] = np.array([n, chi, L, S, seed])
    out[key + "_gates"] = g
    out[key + "_kinds"] = k
    out[key + "_fidelity"] = np.array(O.circuit_fidelity(psi, res["layers"], n))
    out[key + "_state"] = O.circuit_state(res["layers"], n)
    out[key + "_tt_spectrum0"] = rec["tt_svd"][0]
    out[key + "_tt_spectrum_mid"] = rec["tt_svd"][n // 2 - 1]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_canonical.npz"), **out)
print("written", sorted(out))
