"""Config 5 over NCCL (-m gpu, needs >= 2 GPUs: skipped on a one-GPU box): two ranks, states sharded round-robin,
captured-graph lanes + one batched sweeps launch per rank, ONE all_gather_into_tensor of the records; every rank must
hold every record and each must equal the single-GPU eager result of the same state."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, states, cfg, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from qmprs_b200 import batch, host
    from qmprs_b200.kernels import get_kernels
    K = get_kernels(f"cuda:{rank}")
    n, chi, L, S = cfg
    recs = batch.prepare_state_batch(states, chi, L, S, kernels=K, graph_lanes=2)
    mine = batch.shard_indices(len(states), rank, world)
    eager = {s: host.prepare(K, states[s], n, chi, L, S) for s in mine}
    q.put((rank, [(r["n_layers"], r["fidelity"], r["gates"], r["kinds"]) for r in recs],
           {s: (e["n_layers"], e["fidelity"], e["gates"], e["kinds"]) for s, e in eager.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_nccl_gather_matches_eager():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import qmprs_oracle as O
    n, chi, L, S, B = 8, 16, 2, 2, 7
    states = np.stack([O.random_state(n, 500 + s) for s in range(B)])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, states, (n, chi, L, S), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=500) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    eager = {}
    for _, _, e in got:
        eager.update(e)
    assert sorted(eager) == list(range(B))
    for rank, recs, _ in got:                               # every rank holds every record after the gather
        assert len(recs) == B
        for s in range(B):
            L_, f, g, kinds = recs[s]
            eL, ef, eg, ek = eager[s]
            assert L_ == eL and kinds == ek
            assert abs(f - ef) <= 1e-9 and np.abs(g - eg).max() <= 1e-7
            ref = O.prepare(states[s], n, chi, L, S, gauge="canonical")
            assert abs(f - O.circuit_fidelity(states[s], ref["layers"], n)) <= 1e-6
