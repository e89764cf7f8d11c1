#!/usr/bin/env python
"""Benchmark of the qmprs MPS hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c1|c2|c3|c5]

BASELINE.json's metric has two parts and `--workload auto` (default) picks by N:
  N = 1   "prepare_state s/state at 20q chi=512 15 layers": Sequential.prepare_state at the headline
          configuration C3 (20 qubits, chi=512, 15 layers, 50 sweeps).  A step = one prepare_state of a fresh
          synthetic random state (reference distribution, README.md:52-53).  The same line carries, under
          `batch_c5`, one pass of the sharded batch workload on this one GPU: the N = 1 point of the
          multi-GPU curve.
  N > 1   "batch states/sec 1-8 GPU": configuration C5, 4096 random 12-qubit states (chi=64, 10 layers,
          20 sweeps) sharded by state over the N ranks, ending in the path's one collective
          (all_gather_into_tensor of the gate records over NCCL).  A step = the whole batch; strong scaling.
Prints ONE JSON line (task contract): `value` = whole-job states/s with the inputs resident in HBM, `e2e` =
the same through the public call with host buffers, `roofline` for the dominant kernel (timed live with CUDA
events in one extra instrumented step), `cpu_baseline` = the numpy oracle on the host cores on a bounded
sample (N = 1 only).  `--impl reference` times the CPU path only (rank 0).
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


def _cpu_batch_worker(job):
    """One host core: `verbatim` oracle on a few states of the batch, BLAS single-threaded."""
    seeds, n, chi, L, S = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import qmprs_oracle as O
    t0 = time.perf_counter()
    for sd in seeds:
        O.prepare(rand_state(n, sd), n, chi, L, S, gauge="verbatim")
    return time.perf_counter() - t0


class CpuBatchPool:
    """CPU arm of the batch workload: the states are independent, so the host runs one oracle process per
    core (what a user of the reference would do with multiprocessing).  The pool is started (spawn: nothing
    of this process's CUDA state is inherited) and warmed once; every sample times `per_core` states per core."""

    def __init__(self, wl):
        import concurrent.futures as cf
        import multiprocessing as mp
        self.wl = wl
        self.cores = os.cpu_count() or 1
        keep = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
        for k in keep:
            os.environ[k] = "1"                            # inherited by the workers: single-threaded BLAS each
        try:
            self.ex = cf.ProcessPoolExecutor(max_workers=self.cores, mp_context=mp.get_context("spawn"))
            cfg = (wl["n"], wl["chi"], wl["layers"], wl["sweeps"])
            list(self.ex.map(_cpu_batch_worker, [([0], *cfg)] * self.cores))       # start-up + imports, untimed
        finally:
            for k, v in keep.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def sample(self, per_core=2, first_seed=0):
        """Returns (states/s, cores, states timed, wall seconds)."""
        wl = self.wl
        cfg = (wl["n"], wl["chi"], wl["layers"], wl["sweeps"])
        jobs = [([first_seed + c * per_core + j for j in range(per_core)], *cfg) for c in range(self.cores)]
        t0 = time.perf_counter()
        list(self.ex.map(_cpu_batch_worker, jobs))
        wall = time.perf_counter() - t0
        nst = self.cores * per_core
        return nst / wall, self.cores, nst, wall

    def close(self):
        self.ex.shutdown()


def run_reference_batch(args, wl):
    vals = []
    pool = CpuBatchPool(wl)
    for i in range(args.warmup + args.steps):
        v, cores, nst, wall = pool.sample(per_core=2, first_seed=1000 * (i + 1))
        if i >= args.warmup:
            vals.append(v)
    pool.close()
    val = float(np.mean(vals))
    B = wl["batch"]
    sample = (f"per step: {nst} states of the batch, one numpy-oracle process per host core ({cores} cores, BLAS "
              f"single-threaded inside each), {wall:.1f} s wall; value = states timed / wall (the full batch of {B} is "
              f"{B / val:.0f} s at this rate)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": B / val * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic", "config": batch_config(wl, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "measured_sample_s": wall,
    }))


def batch_config(wl, world):
    return {"workload": wl["name"], "n_qubits": wl["n"], "chi": wl["chi"], "layers": wl["layers"],
            "sweeps": wl["sweeps"], "batch": wl["batch"]}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "batch" in wl:
        return run_reference_batch(args, wl)
    use_all_host_threads()
    setup = None
    totals = []
    for i in range(args.warmup + args.steps):
        t_s = time.perf_counter()
        tot, br, setup = cpu_sample(wl, 1000 + i, setup)
        t_s = time.perf_counter() - t_s
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
        # wall time of one step's bounded sample (ms_per_step is the EXTRAPOLATED full-workload time: it does
        # not fit inside this process's run time by construction)
        "measured_sample_s": t_s,
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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


SVD_STAGE = ("svd_gram", "svd_eig", "svd_apply", "svd_layout", "svd_round")


def roofline_from_profile(prof, peak_tf):
    """`prof`: {class: (ms, launches, algorithmic work)} of one instrumented step (CUDA events around every
    launch).  Reports the TIME-DOMINANT part of the step.  A Jacobi round is Gram -> 32x32 eigen-solve -> update
    (one cooperative k_round launch per outer sweep; three kernels per round on the QM_SVD_SCHED=grouped path), of
    which the eigen-solve is latency bound and carries no flops of its own, so the SVD is reported as a stage:
    algorithmic flops of its Gram + update GEMMs over the time of all its kernels.  Streaming classes are reported against the measured HBM bandwidth."""
    peaks = load_peaks()
    tot_ms = sum(v[0] for v in prof.values())
    svd_ms = sum(prof[k][0] for k in SVD_STAGE if k in prof)
    svd_work = sum(prof[k][2] for k in SVD_STAGE if k in prof and k != "svd_layout")
    groups = {"svd": (svd_ms, sum(prof[k][1] for k in SVD_STAGE if k in prof), svd_work)}
    for k, v in prof.items():
        if k not in SVD_STAGE:
            groups[k] = v
    dom = max(groups, key=lambda k: groups[k][0])
    d_ms, d_cnt, d_work = groups[dom]
    if dom in ("gate", "env_polar"):
        hbm = peaks.get("hbm_gbs", 6650.0)
        ach = d_work / (d_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"}
    else:
        ach = d_work / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": None,
                "peak_source": "FP64: cuBLAS ZGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}
    name = {"svd": "svd stage: k_round (fused Gram + eigen-solve + update per Jacobi round) + layout", "zgemm": "k_zgemm_tma",
            "env_polar": "k_env_fused", "gate": "k_gate2"}.get(dom, dom)
    try:      # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, cold cache)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        t = tj.get("svd_round" if dom == "svd" else dom)
        roof["traffic"] = t["bytes"] if t else None
        roof["traffic_note"] = t.get("note") if t else None
    except Exception:
        pass
    def tf(k):
        ms, cnt, work = prof[k]
        return {"ms": round(ms, 3), "launches": cnt, "share": ms / tot_ms, "avg_launch_us": 1e3 * ms / max(cnt, 1),
                "work": work, "rate": (work / (ms * 1e-3) / (1e9 if k in ("gate", "env_polar") else 1e12)) if ms > 0 and work > 0 else None}
    roof.update({"kernel": name, "launches": d_cnt, "avg_launch_us": 1e3 * d_ms / max(d_cnt, 1),
                 "share_of_kernel_time": d_ms / tot_ms,
                 "classes": {k: tf(k) for k in prof if prof[k][0] > 0.0}})
    return roof


def bench_env(args):
    import torch
    import torch.distributed as dist
    from qmprs_b200.kernels import get_kernels
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K = get_kernels(str(dev))

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

    return dict(torch=torch, dist=dist, world=world, rank=rank, local=local, dev=dev, K=K, sync_all=sync_all,
                max_over_ranks=max_over_ranks)


def run_ours(args, wl):
    env = bench_env(args)
    torch, dist, world, rank, dev, K = env["torch"], env["dist"], env["world"], env["rank"], env["dev"], env["K"]
    if "batch" in wl:
        line = measure_batch(args, wl, env, args.steps, args.warmup, main_line=True)
    else:
        line = measure_single(args, wl, env)
        if rank == 0 and args.workload == "auto" and world == 1:
            # N = 1 point of the multi-GPU curve: one pass of the sharded batch workload on this GPU
            try:
                b = measure_batch(args, WORKLOADS["c5"], env, steps=1, warmup=1, main_line=False)
                line["batch_c5"] = {k: b[k] for k in ("value", "unit", "ms_per_step", "config", "e2e", "gpu_launches",
                                                      "graph", "fidelity_mean", "cpu_baseline", "scaling")}
                try:
                    json.dump({"value": b["value"], "e2e": b["e2e"]["value"], "when": time.time()},
                              open(C5_N1_FILE, "w"))
                except Exception:
                    pass
            except Exception as ex:                       # never lose the headline line to the extra pass
                line["batch_c5"] = {"error": repr(ex)}
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


C5_N1_FILE = "/tmp/qmprs_b200_bench_c5_n1.json"


def measure_single(args, wl, env):
    """One state per step per GPU (C1-C3): `value` with the state resident in HBM, `e2e` through
    Sequential.prepare_state with a pinned host buffer in and the gate records out."""
    torch, world, rank, local, dev, K = env["torch"], env["world"], env["rank"], env["local"], env["dev"], env["K"]
    sync_all, max_over_ranks = env["sync_all"], env["max_over_ranks"]
    from qmprs_b200 import GateListCircuit, host
    from qmprs.synthesis.mps_encoding import Sequential
    n, chi, L, S = wl["n"], wl["chi"], wl["layers"], wl["sweeps"]
    W, Ksteps = args.warmup, args.steps
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
    enc.prepare_state(states[0], chi, num_layers=L, num_sweeps=S)
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
    if rank != 0:
        return None

    # ---- roofline: one extra instrumented step (CUDA events per launch) ----
    peak_tf = measure_fp64_peak(torch, dev)
    K.prof_begin()
    host.prepare(K, dev_states[-1], n, chi, L, S, split=args.split)
    roof = roofline_from_profile(K.prof_end(), peak_tf)
    # same workload with the opt-in SVD-free re-split (identical circuit), reported beside `value`
    alt = None
    if args.split == "svd":
        host.prepare(K, dev_states[0], n, chi, L, S, split="exact")
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        na = min(Ksteps, 3)
        falt = [host.prepare(K, dev_states[W + i], n, chi, L, S, split="exact")["fidelity"] for i in range(na)]
        a1.record()
        torch.cuda.synchronize(dev)
        alt = {"split": "exact", "value": na / (a0.elapsed_time(a1) * 1e-3), "unit": UNIT,
               "max_abs_fidelity_diff_vs_svd": float(np.max(np.abs(np.array(falt) - np.array(fid[W:W + na]))))}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        tot, br, _ = cpu_sample(wl, 4242)
        cpu = {"value": 1.0 / tot, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": ("numpy oracle at full size: MPS build + 1 disentangling layer + 1 layer of sweep gate-steps, "
                          f"extrapolated t_mps + L*t_layer + S*L*t_sweep_layer = {br['t_mps']:.2f} + {L}*{br['t_layer']:.2f}"
                          f" + {S}*{L}*{br['t_sweep_layer']:.2f} s")}
    return {
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


def measure_batch(args, wl, env, steps, warmup, main_line):
    """Config 5: a step = the whole batch of independent states sharded over the ranks (state s -> rank
    s mod world), every state one CUDA-graph replay on one of `--lanes` concurrent lanes, followed by the
    single all-gather of the gate records.  `value`: the rank's shard of the states is resident in HBM when
    the timed region starts and the gathered records stay in HBM; `e2e`: numpy states in, list of records out
    through qmprs_b200.batch.prepare_state_batch (pinned H2D of the shard, D2H of the gathered records)."""
    torch, world, rank, local, dev, K = env["torch"], env["world"], env["rank"], env["local"], env["dev"], env["K"]
    sync_all, max_over_ranks = env["sync_all"], env["max_over_ranks"]
    from qmprs_b200 import batch as qb
    from qmprs_b200 import host
    from qmprs_b200.graphs import GraphedPreparer
    n, chi, L, S = wl["n"], wl["chi"], wl["layers"], wl["sweeps"]
    B = args.batch or wl["batch"]
    states = np.stack([rand_state(n, s) for s in range(B)])          # seed = state index
    prep = GraphedPreparer(n, chi, L, S, lanes=args.lanes, device=str(dev), width=args.width) if args.lanes > 0 else None
    sdev = torch.from_numpy(states).to(dev)
    small = 2 * world * max(args.lanes, 1)
    for w in range(max(warmup, 1)):
        # the first warm-up pass is a short one (lazy initialisation), later ones the full batch
        qb.prepare_state_batch(sdev if w > 0 else sdev[:small], chi, L, S, kernels=K, preparer=prep, return_device=True)
    clocks = ClockSampler(local)
    sync_all()
    clocks.start()
    l0, r0 = K.launch_count(), (prep.replays if prep else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        recs_dev = qb.prepare_state_batch(sdev, chi, L, S, kernels=K, preparer=prep, return_device=True)
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    launches = int(K.launch_count() - l0 + ((prep.replays - r0) * prep.nodes_per_graph if prep else 0))
    value = steps * B / (ms * 1e-3)
    # ---- end to end: host states in, host records out ----
    # two untimed passes first: the page-locked staging blocks (input shard; two result blocks, one still referenced
    # by the previous step's records while the next is filled) are allocated once and recycled afterwards
    for _ in range(2 if warmup > 0 else 0):
        recs = qb.prepare_state_batch(states, chi, L, S, kernels=K, preparer=prep)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        recs = qb.prepare_state_batch(states, chi, L, S, kernels=K, preparer=prep)
    e1.record()
    sync_all()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = steps * B / (ms_e2e * 1e-3)
    phases = None
    if main_line and prep is not None and prep.defer:
        # one more pass with events between the two phases of the batch path -- on EVERY rank: the pass ends with the
        # all-gather (a collective entered by rank 0 alone would never complete)
        prep.time_phases = True
        qb.prepare_state_batch(sdev, chi, L, S, kernels=K, preparer=prep, return_device=True)
        torch.cuda.synchronize(dev)
        prep.time_phases = False
        phases = prep.phase_ms
    sync_all()
    if rank != 0:
        return None
    rl = qb.record_len(n, L)
    fid = float(np.mean([r["fidelity"] for r in recs]))
    roof = None
    cpu = None
    if main_line:
        # instrumented EAGER pass over a few states (graph replays bypass the per-launch events)
        peak_tf = measure_fp64_peak(torch, dev)
        K.prof_begin()
        for s in range(4):
            host.prepare(K, sdev[s], n, chi, L, S)
        eager = roofline_from_profile(K.prof_end(), peak_tf)
        roof = None
        if phases is not None:
            # the batch path proper (pass above).  Its dominant single kernel is the whole-shard sweeps launch
            # (k_sweeps_small: every sweep of every state of the rank's shard, one CTA per state, vectors in shared
            # memory); algorithmic bytes = those of the unfused kernels it replaces, 96 B per amplitude and gate-step
            # (32 forward + 64 backward), so "achieved" is the HBM traffic the launch AVOIDS per second.
            layers_ms, sweeps_ms = phases
            peaks = load_peaks()
            hbm = peaks.get("hbm_gbs", 6650.0)
            shard = (sdev.shape[0] + world - 1) // world               # states of this rank (the launch is per rank)
            work = float(shard) * S * (L * n) * 96.0 * float(2 ** n)
            ach = work / (sweeps_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                    "kernel": "k_sweeps_small (all sweeps of a state in one CTA, vectors in shared memory)",
                    "launches": 1, "avg_launch_us": 1e3 * sweeps_ms, "share": sweeps_ms / (layers_ms + sweeps_ms),
                    "phases_ms": {"layer_graphs": layers_ms, "sweeps_launch": sweeps_ms},
                    "note": ("CUDA events between the two phases of one extra pass over this rank's shard; the vectors "
                             "live in shared memory (DRAM traffic ~0: ncu_r02_summary.md), so achieved/peak compares "
                             "the launch with an HBM-streaming implementation of the same gate-steps, it is not a "
                             "DRAM utilisation; the kernel is bound by the one-warp 4x4 polar between its passes "
                             "(profiles/sweeps_small_phases_r02_*.log)"),
                    "eager_classes": eager.get("classes")}
        if roof is None:
            roof = eager
            if roof.get("kernel") == "k_env_fused":
                roof["kernel"] = "k_sweeps_small (all sweeps of a state in one CTA, vectors in shared memory)"
            roof["note"] = ("eager instrumented pass over 4 states; the timed region replays the same kernels from CUDA "
                            "graphs, where launch latency (not any pipe) bounds a 12-qubit state; the sweeps kernel works "
                            "out of shared memory, so its HBM fraction is not a utilisation figure")
    if world == 1 and not args.no_cpu_baseline:
        pool = CpuBatchPool(wl)
        v, cores, nst, wall = pool.sample(per_core=2, first_seed=0)
        pool.close()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"{nst} states of the batch, one numpy-oracle process per host core ({cores} cores, BLAS "
                          f"single-threaded inside each), {wall:.1f} s wall")}
    n1 = None
    if world > 1:
        try:
            j = json.load(open(C5_N1_FILE))
            if time.time() - j["when"] < 4 * 3600:
                n1 = {"value": j["value"], "e2e": j["e2e"], "source": "batch_c5 of this box's preceding --gpus 1 run"}
        except Exception:
            pass
    cfg = batch_config(wl, world)
    cfg.update({"batch": B, "graph_lanes_per_gpu": args.lanes, "l2": "inputs larger than L2: 256 MiB of states per batch",
                "parallelism": f"states sharded over {world} GPU(s) (state s -> rank s mod {world}), "
                               "one all_gather_into_tensor of the records (NCCL)"})
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128", "data": "synthetic", "config": cfg,
        "fidelity_mean": fid, "cpu_baseline": cpu, "clocks": clk, "roofline": roof,
        "gpu_launches": launches,
        "graph": ({"lanes": args.lanes, "states_per_graph": prep.width, "kernel_nodes_per_state": prep.nodes_per_graph,
                   "eager_fallbacks": prep.fallbacks} if prep else None),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * 16 * 2 ** n),
                "d2h_bytes_per_step": int(world * ((B + world - 1) // world) * rl * 8)},
        "n1_same_workload": n1,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=32, help="concurrent CUDA-graph lanes per GPU for --workload c5 (0 = eager)")
    ap.add_argument("--width", type=int, default=16, help="states per captured graph (advanced in lock step) in the batch workload")
    ap.add_argument("--batch", type=int, default=0, help="states per step for the batch workload (default: config 5's 4096)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS),
                    help="auto: c3 (headline single state) at --gpus 1, c5 (sharded batch + NCCL gather) at --gpus N > 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split", default="svd", choices=["svd", "exact"],
                    help="two-site re-split: 'svd' = reference arithmetic (default), 'exact' = gauge-free, no SVD")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[("c3" if max(world, args.gpus) == 1 else "c5") if args.workload == "auto" else args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
