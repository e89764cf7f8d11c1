#!/usr/bin/env python
"""Benchmark of the qmprs MPS hot path on B200:  Sequential.prepare_state at the
BASELINE.json headline configuration (20 qubits, chi=512, 15 layers, 50 sweeps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1]

A "step" is one prepare_state of a fresh synthetic random state (reference distribution,
README.md:52-53).  Prints ONE JSON line (see the task contract): `value` = whole-job
states/s with the input resident in HBM, `e2e` = the same through the public
Sequential.prepare_state call with host buffers, `roofline` for the dominant kernel
(timed live with CUDA events in one extra instrumented step), `cpu_baseline` = the numpy
oracle on the host cores on a bounded sample.  `--impl reference` times the CPU path only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per graph lane (the default of 8 serialises lanes that share a queue: 109 -> 380
# twelve-qubit states/s on one B200); read by the driver when the CUDA context is created
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOADS = {
    "c1": dict(n=10, chi=512, layers=15, sweeps=50, name="10q chi=512(32) 15 layers 50 sweeps (README)"),
    "c2": dict(n=16, chi=256, layers=15, sweeps=50, name="16q chi=256 15 layers 50 sweeps"),
    "c3": dict(n=20, chi=512, layers=15, sweeps=50, name="20q chi=512 15 layers 50 sweeps (headline)"),
}
WORKLOADS["c5"] = dict(n=12, chi=64, layers=10, sweeps=20, batch=4096,
                       name="batch of 12q states chi=64 10 layers 20 sweeps, sharded by state + NCCL gather")
METRIC = "prepare_state_throughput"
UNIT = "states/s"


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.random(2 ** n) + 1j * rng.random(2 ** n)
    return v / np.linalg.norm(v)


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores, bounded sample extrapolated to the full workload
# ----------------------------------------------------------------------------------------
def cpu_sample(wl, seed, setup=None):
    """Times one disentangling layer and one layer's worth (N gates) of sweep gate-steps of
    the verbatim oracle on the real workload shapes; the MPS build is timed once in
    ``setup``.  Returns (seconds per state extrapolated, breakdown, setup)."""
    from oracle import qmprs_oracle as O
    n, chi, L, S = wl["n"], wl["chi"], wl["layers"], wl["sweeps"]
    if setup is None:
        psi = rand_state(n, seed)
        t0 = time.perf_counter()
        A = O.compress_right(O.from_dense(psi, n), max_bond=chi)
        target = O.to_dense(A)
        B = [a.copy() for a in A]
        nrm = O.mps_norm(B)
        if not np.isclose(nrm, 1.0):
            B[-1] = B[-1] / nrm
        B = O.right_canon(O.compress_right(B), normalize=True)
        t_mps = time.perf_counter() - t0
        setup = dict(B=B, target=target, t_mps=t_mps)
    B = [b.copy() for b in setup["B"]]
    t0 = time.perf_counter()
    C = O.chi2_truncate(B, "verbatim")
    layer = O.generate_unitary_layer(C, "verbatim")
    O.apply_inverse_layer(B, layer)
    O.zero_overlap(B)
    t_layer = time.perf_counter() - t0
    setup["B"] = B                      # next sample continues from the disentangled MPS
    layers = [layer]
    t0 = time.perf_counter()
    O.sweep(setup["target"], layers, n)   # builds the 1-layer circuit state + N gate-steps
    t_sweep_layer = time.perf_counter() - t0
    total = setup["t_mps"] + L * t_layer + S * L * t_sweep_layer
    return total, dict(t_mps=setup["t_mps"], t_layer=t_layer, t_sweep_layer=t_sweep_layer), setup


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """The CPU arm runs with every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    leave the BLAS behind numpy single-threaded; raise the pools back to the core count."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    setup = None
    totals = []
    for i in range(args.warmup + args.steps):
        tot, br, setup = cpu_sample(wl, 1000 + i, setup)
        if i >= args.warmup:
            totals.append(tot)
    sec = float(np.mean(totals))
    val = 1.0 / sec
    sample = ("per step: 1 disentangling layer + 1 layer of sweep gate-steps of the numpy oracle at full size; "
              "MPS build timed once; extrapolated t_mps + L*t_layer + S*L*t_sweep_layer "
              f"(last: {br['t_mps']:.2f}s, {br['t_layer']:.2f}s, {br['t_sweep_layer']:.2f}s); "
              "the reference's per-sweep from_dense (sequential.py:443) is not charged")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": wl["name"], "n_qubits": wl["n"], "chi": wl["chi"], "layers": wl["layers"],
                   "sweeps": wl["sweeps"]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, dev):
    """cuBLAS ZGEMM 4096^3 burst (8*n^3 real flops), best of 5: the FP64 roofline denominator."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device=dev)
    b = torch.randn(n, n, dtype=torch.complex128, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 8.0 * n ** 3 / (best * 1e-3) / 1e12


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from qmprs_b200 import GateListCircuit, host
    from qmprs_b200.kernels import get_kernels
    from qmprs.synthesis.mps_encoding import Sequential

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    K = get_kernels(str(dev))
    n, chi, L, S = wl["n"], wl["chi"], wl["layers"], wl["sweeps"]
    W, Ksteps = args.warmup, args.steps

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if "batch" in wl:
        return run_batch(args, wl, K, dev, world, rank, sync_all, max_over_ranks)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    states = [rand_state(n, rank * 100003 + i) for i in range(W + Ksteps)]
    dev_states = [K.from_host(s) for s in states]

    # ---- device-resident value ----
    fid = []
    for i in range(W):
        flush.fill_(i)
        fid.append(host.prepare(K, dev_states[i], n, chi, L, S, split=args.split)["fidelity"])
    clocks = ClockSampler(local)
    sync_all()
    clocks.start()
    l0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + Ksteps):
        flush.fill_(i)
        fid.append(host.prepare(K, dev_states[i], n, chi, L, S, split=args.split)["fidelity"])
    e1.record()
    sync_all()
    launches = K.launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    value = world * Ksteps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ----
    enc = Sequential(GateListCircuit)
    enc.gate_split = args.split
    pinned = [torch.from_numpy(s).pin_memory().numpy() for s in states[W:]]
    for s in states[:2]:      # untimed: lets small registers (n <= 16) switch to their captured graph (steady state)
        enc.prepare_state(s, chi, num_layers=L, num_sweeps=S)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2h = 0
    for p in pinned:
        flush.fill_(1)
        circ = enc.prepare_state(p, chi, num_layers=L, num_sweeps=S)
        d2h = sum(m.nbytes for m, _ in circ.gates)
    e1.record()
    sync_all()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = world * Ksteps / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel: one extra instrumented step (CUDA events per launch) ----
        peak_tf = measure_fp64_peak(torch, dev)
        K.prof_begin()
        host.prepare(K, dev_states[-1], n, chi, L, S, split=args.split)
        prof = K.prof_end()
        # same workload with the opt-in SVD-free re-split (identical circuit), reported beside `value`
        alt = None
        if args.split == "svd":
            host.prepare(K, dev_states[0], n, chi, L, S, split="exact")
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            falt = [host.prepare(K, dev_states[W + i], n, chi, L, S, split="exact")["fidelity"] for i in range(Ksteps)]
            a1.record()
            torch.cuda.synchronize(dev)
            alt = {"split": "exact", "value": Ksteps / (a0.elapsed_time(a1) * 1e-3), "unit": UNIT,
                   "max_abs_fidelity_diff_vs_svd": float(np.max(np.abs(np.array(falt) - np.array(fid[W:W + Ksteps]))))}
        tot_ms = sum(v[0] for v in prof.values())
        # dominant kernel with a throughput roofline; latency-bound classes (no algorithmic-work model:
        # the 32x32 shared-memory eigen-solve, per-column Householder vectors, small kernels) are reported
        # as time shares only ("latency/occupancy-bound, report as time only")
        latency = {k: {"share": v[0] / tot_ms, "avg_launch_us": 1e3 * v[0] / max(v[1], 1)}
                   for k, v in prof.items() if v[2] == 0.0 and v[0] > 0.0}
        dom = max((k for k in prof if prof[k][2] > 0.0), key=lambda k: prof[k][0])
        d_ms, d_cnt, d_work = prof[dom]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if dom in ("gate", "env_polar"):
            hbm = peaks.get("hbm_gbs", 6650.0)
            ach = d_work / (d_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": None, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}
        else:
            ach = d_work / (d_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "traffic": None,
                    "peak_source": "FP64: cuBLAS ZGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, cold cache)
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
            roof["traffic"] = traffic["bytes"] if traffic else None
            roof["traffic_note"] = traffic.get("note") if traffic else None
        except Exception:
            pass
        roof.update({"kernel": dom, "launches": d_cnt, "avg_launch_us": 1e3 * d_ms / max(d_cnt, 1),
                     "share_of_kernel_time": d_ms / tot_ms, "latency_bound_classes": latency,
                     "classes": {k: {"ms": round(v[0], 3), "launches": v[1], "work": v[2]} for k, v in prof.items()}})
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            use_all_host_threads()
            tot, br, _ = cpu_sample(wl, 4242)
            cpu = {"value": 1.0 / tot, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                   "sample": ("numpy oracle at full size: MPS build + 1 disentangling layer + 1 layer of sweep gate-steps, "
                              f"extrapolated t_mps + L*t_layer + S*L*t_sweep_layer = {br['t_mps']:.2f} + {L}*{br['t_layer']:.2f}"
                              f" + {S}*{L}*{br['t_sweep_layer']:.2f} s")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Ksteps, "warmup": W,
            "ms_per_step": ms / Ksteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": wl["name"], "n_qubits": n, "chi": chi, "layers": L, "sweeps": S,
                       "states_per_step_per_gpu": 1, "gate_split": args.split, "l2": "256 MiB buffer rewritten between steps",
                       "parallelism": f"{world} independent states (one per GPU), no data-path collective"},
            "s_per_state": ms * 1e-3 / Ksteps,
            "fidelity_mean": float(np.mean(fid[W:])),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(16 * 2 ** n), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "alt_exact_split": alt,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_batch(args, wl, K, dev, world, rank, sync_all, max_over_ranks):
    """Config 5: a step = one batch of `--batch` independent states sharded over the ranks
    (state s -> rank s mod world) followed by the single all-gather of the gate records."""
    import torch
    import torch.distributed as dist
    from qmprs_b200 import batch as qb
    n, chi, L, S = wl["n"], wl["chi"], wl["layers"], wl["sweeps"]
    B = args.batch
    from qmprs_b200.graphs import GraphedPreparer
    states = np.stack([rand_state(n, s) for s in range(B)])          # seed = state index
    prep = GraphedPreparer(n, chi, L, S, lanes=args.lanes, device=str(dev)) if args.lanes > 0 else None
    for _ in range(max(args.warmup, 1)):
        qb.prepare_state_batch(states[: 2 * world * max(args.lanes, 1)], chi, L, S, kernels=K, preparer=prep)
    sync_all()
    l0 = K.launch_count()
    r0 = prep.replays if prep else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        recs = qb.prepare_state_batch(states, chi, L, S, kernels=K, preparer=prep)
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = args.steps * B / (ms * 1e-3)
    if rank == 0:
        fid = float(np.mean([r["fidelity"] for r in recs]))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import qmprs_oracle as O
            use_all_host_threads()
            t0 = time.perf_counter()
            nref = 4
            for s in range(nref):
                O.prepare(states[s], n, chi, L, S, gauge="verbatim")
            dt = (time.perf_counter() - t0) / nref
            cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                   "sample": f"numpy oracle, {nref} states of the batch run one after the other (BLAS threads as configured)"}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": wl["name"], "n_qubits": n, "chi": chi, "layers": L, "sweeps": S, "batch": B,
                       "parallelism": f"states sharded over {world} GPU(s), one all-gather of records"},
            "fidelity_mean": fid, "cpu_baseline": cpu,
            "gpu_launches": int(K.launch_count() - l0 + ((prep.replays - r0) * prep.nodes_per_graph if prep else 0)),
            "graph": ({"lanes": args.lanes, "kernel_nodes_per_state": prep.nodes_per_graph,
                       "eager_fallbacks": prep.fallbacks} if prep else None),
            "value_note": "the batch workload is timed end to end only: host states in (pinned H2D per state), host gate records out",
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(B * 16 * 2 ** n // world),
                    "d2h_bytes_per_step": int(B * qb.record_len(n, L) * 8)},
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=32, help="concurrent CUDA-graph lanes per GPU for --workload c5 (0 = eager)")
    ap.add_argument("--batch", type=int, default=64, help="states per step for --workload c5 (config 5 uses 4096)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split", default="svd", choices=["svd", "exact"],
                    help="two-site re-split: 'svd' = reference arithmetic (default), 'exact' = gauge-free, no SVD")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
