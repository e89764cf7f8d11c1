"""Writes tests/golden/oracle_schedules.npz: the Iter DiOall / Iter DiOi schedules of the oracle (canonical gauge)
on two small registers.  Like oracle_canonical.npz these fixtures pin the ORACLE (the reference names the schedules
but does not implement them: notebook :459, sequential.py:410, 428-432).
Regenerate with:  python tests/golden/make_golden_schedules.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qmprs_oracle as O  # noqa: E402

CASES = {"s8": (8, 32, 4, 3, 2), "s10": (10, 32, 3, 2, 5)}
out = {}
for key, (n, chi, L, S, seed) in CASES.items():
    psi = O.random_state(n, seed)
    for schedule in ("IterDiOall", "IterDiOi"):
        res = O.prepare(psi, n, chi, L, S, gauge="canonical", schedule=schedule)
        g = np.zeros((L, n, 16), dtype=complex)
        for li, _, _, site, G in O.flatten_layers(res["layers"]):
            g[li, site, : G.size] = G.reshape(-1)
        tag = f"{key}_{schedule}"
        out[tag + "_cfg"] = np.array([n, chi, L, S, seed])
        out[tag + "_gates"] = g
        out[tag + "_fidelity"] = np.array(O.circuit_fidelity(psi, res["layers"], n))
        out[tag + "_overlaps"] = np.array(res["overlaps"])
        out[tag + "_state"] = O.circuit_state(res["layers"], n)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_schedules.npz"), **out)
print("written", sorted(out))
