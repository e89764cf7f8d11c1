"""N>1 path on CPU: world_size-2 gloo process group, states sharded round-robin, one
all-gather of the records.  Uses the numpy test double for the
kernels; the NCCL run of the same code is bench.py --workload c5 --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qmprs_oracle as O
from qmprs_b200 import batch


def test_shard_and_record_roundtrip():
    assert batch.shard_indices(7, 1, 3) == [1, 4]
    assert sorted(sum((batch.shard_indices(10, r, 4) for r in range(4)), [])) == list(range(10))
    res = {"gates": np.arange(2 * 3 * 16).reshape(2, 3, 16) * (1 + 2j), "kinds": [[2, 2, 1], [2, 1, 1]],
           "n_layers": 2, "fidelity": 1.25, "overlap": (0.75, -1.0)}
    vec = batch.pack_record(res, 3, 4)
    assert vec.shape == (batch.record_len(3, 4),)
    back = batch.unpack_record(vec, 3, 4)
    assert back["n_layers"] == 2 and back["kinds"] == res["kinds"] and back["fidelity"] == 1.25 and back["overlap"] == (0.75, -1.0)
    assert np.array_equal(back["gates"], res["gates"])


def test_unpack_records_bulk_ragged_layers():
    """Bulk unpacking of a (B, record_len) block: ragged layer counts, odd record length (views into unaligned
    rows), same content as the per-record form."""
    rng = np.random.default_rng(0)
    n, L = 3, 4
    recs = []
    for nl in (4, 1, 3, 2, 4):
        recs.append({"gates": rng.standard_normal((nl, n, 16)) + 1j * rng.standard_normal((nl, n, 16)),
                     "kinds": [[2, 2, 1]] * nl, "n_layers": nl, "fidelity": 0.0,
                     "overlap": (float(rng.standard_normal()), float(rng.standard_normal()))})
    block = np.stack([batch.pack_record(r, n, L) for r in recs])
    assert block.shape[1] % 2 == 1                       # odd row length: every second row starts on an 8-byte boundary
    out = batch.unpack_records(block, n, L)
    assert len(out) == len(recs)
    for got, ref in zip(out, recs):
        assert got["n_layers"] == ref["n_layers"] and got["kinds"] == ref["kinds"] and got["overlap"] == ref["overlap"]
        assert got["gates"].shape == ref["gates"].shape and np.array_equal(got["gates"], ref["gates"])
        assert abs(got["fidelity"] - np.hypot(*ref["overlap"])) < 1e-15
        assert np.array_equal(np.conj(got["gates"]).T @ np.ones(got["gates"].shape[0]), np.conj(ref["gates"]).T @ np.ones(ref["gates"].shape[0]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, states, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.fake_kernels import FakeKernels
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = batch.prepare_state_batch(states, 16, num_layers=2, num_sweeps=1, kernels=FakeKernels())
    q.put((rank, [(r["n_layers"], r["fidelity"], r["gates"]) for r in out]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_gather_matches_oracle():
    n, B = 5, 5
    states = np.stack([O.random_state(n, s) for s in range(B)])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, states, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for s in range(B):
        ref = O.prepare(states[s], n, 16, 2, 1, gauge="canonical")
        fref = O.circuit_fidelity(states[s], ref["layers"], n)
        for rank in (0, 1):                       # every rank holds every record after the gather
            L, f, g = got[rank][s]
            assert L == ref["n_layers"] and abs(f - fref) < 1e-9
            for li, _, _, site, G in O.flatten_layers(ref["layers"]):
                assert np.abs(g[li, site, : G.size] - G.reshape(-1)).max() < 1e-8
