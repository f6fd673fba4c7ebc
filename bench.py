#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the B200 engine.

Metric (BASELINE.json): marker updates per second of the ppbayes Gibbs sweep (and seconds per sweep)
on synthetic dense 2-bit genotypes N=500k x M=1M; achieved HBM GB/s vs the measured peak.

A "step" is ONE full Gibbs sweep (= one AnalysisGraph::exec: M marker updates in a fresh random order,
each = x_j.eps dot, mixture/beta draw, residual update).  The chain between sweeps (mu, sigmaE, sigmaG,
pi draws) runs for real on the host, exactly as BayesRBase::runGibbs does, so the timed sweeps are
sweeps 1+W .. W+K of a genuine chain started from the reference's initial state.

    value  = M*K / (sum of the K sweep-kernel durations, CUDA events on the engine's stream; inputs staged in HBM)
    e2e    = M*K / (sum of the K brr_gpu_sweep() wall times: pinned-host order + (u,z) table H2D, kernel,
                    D2H of v/betasqnG/scalars and of beta, components, acum -- what exec() must leave in the Analysis)
    N > 1:  individuals sharded over ranks (strong scaling: the total problem is fixed), max over ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
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

SEED = 20260101
# ONE metric string for both arms (the driver pairs the lines by it); sizes live in `config`.
METRIC = "marker updates/sec (ppbayes Gibbs sweep, dense)"
ARITHMETIC = ("fp64 chain state (eps, beta, draws); x_j.eps as exact 64-bit fixed-point integer dots (eps quantised per sweep to "
              "2^-s with >= 57 significant bits relative to max|eps|, 8 int8 limbs x 2-bit codes on integer MMA, int64 reductions)")
MAF_LO, MAF_HI, P_MISSING = 0.01, 0.5, 0.001
S_CVA = np.array([[0.01, 0.001, 0.0001]])
N_CAUSAL, H2 = 1000, 0.5
V0E = S02E = V0G = S02G = 1e-4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-inds", type=int, default=int(os.environ.get("BRR_BENCH_N", 500000)))
    ap.add_argument("--n-markers", type=int, default=int(os.environ.get("BRR_BENCH_M", 1000000)))
    ap.add_argument("--block", type=int, default=int(os.environ.get("BRR_BENCH_BLOCK", 0)))
    ap.add_argument("--cpu-markers", type=int, default=int(os.environ.get("BRR_BENCH_CPU_M", 384)))
    ap.add_argument("--cpu-sweeps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def inv_scaled_chisq(rng, dof, scale):
    # Distributions_boost::inv_scaled_chisq_rng (reference src/distributions_boost.cpp:45-60)
    return 1.0 / rng.gamma(0.5 * dof, 1.0 / (0.5 * dof * scale))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_phenotype(n_inds, n_markers):
    """The bench phenotype rebuilt WITHOUT the product library (reference arm): the same rng stream as the engine arm
    (causal markers, effects, noise), the causal columns from the numpy twin of the synthetic generator, standardised by the
    reference's own DenseMarkerBuilder (oracle/_ref), y = noise + sum_j b_j x_j, then BayesRBase::init's standardisation."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_lib as R
    import synth_np
    rng = np.random.default_rng(SEED)
    causal = rng.choice(n_markers, N_CAUSAL, replace=False)
    bcausal = rng.standard_normal(N_CAUSAL) * np.sqrt(H2 / N_CAUSAL)
    y = rng.standard_normal(n_inds) * np.sqrt(1.0 - H2)
    for j, b in zip(causal, bcausal):
        col = synth_np.synth_bed(SEED, n_inds, [int(j)], MAF_LO, MAF_HI, False, P_MISSING)[0]
        y += b * R.dense_column(col, n_inds)
    y = y.astype(np.float32).astype(np.float64)
    y -= y.sum() / n_inds
    return y / np.sqrt((y ** 2).sum() / (n_inds - 1))


def cpu_reference_leg(n_inds, n_markers, eps_y, steps, warmup=0):
    """The reference's own CPU implementation of the path (oracle/_ref, built from /root/reference with the
    shipped -O3 -ffast-math flags), on a bounded sample: the first `n_markers` markers of the SAME synthetic
    data set, all N individuals, dense layout, marker cache (Sequential graph: 1 core by construction).
    No product library is mapped by this leg: the data set comes from oracle/synth_np.py (numpy twin of the device generator)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_lib as R
    import synth_np
    fast = R.ref_available(fast=True)
    if not fast and not R.ref_available():
        return None
    bed = synth_np.synth_bed(SEED, n_inds, np.arange(n_markers), MAF_LO, MAF_HI, False, P_MISSING)
    y = np.asarray(eps_y, dtype=np.float32)
    _, t = R.run_chain(bed, n_inds, n_markers, y, S_CVA, seed=1, iterations=warmup + steps, layout=0, dump=False, fast=fast)
    per_sweep = t["total_s"] / (warmup + steps)
    # the reference's all-core path: asyncppbayes under its own TBB ParallelGraph (a DIFFERENT chain: bounded staleness of 20 markers)
    all_cores = None
    try:
        _, ta = R.run_chain(bed, n_inds, n_markers, y, S_CVA, seed=1, iterations=max(1, steps), layout=0, dump=False, fast=fast,
                            graph="parallel", tokens=20)
        pa = ta["total_s"] / max(1, steps)
        all_cores = {"value": n_markers / pa, "unit": "marker updates/s", "cores": os.cpu_count(), "kind": "reference",
                     "sample": "reference asyncppbayes (ParallelGraph, TBB, 20 analysis tokens, marker cache) on the same sample, %d sweeps, "
                               "%.3f s/sweep" % (max(1, steps), pa)}
    except Exception as ex:
        all_cores = {"error": repr(ex)}
    return {"value": n_markers / per_sweep, "unit": "marker updates/s", "cores": 1, "kind": "reference", "all_cores_async": all_cores,
            "sample": "reference BayesRBase::runGibbs (dense, marker cache, Sequential graph), N=%d individuals x first %d "
                      "markers of the bench data set, %d sweeps, %.3f s/sweep; linear in M: %.0f s per 1M-marker sweep"
                      % (n_inds, n_markers, warmup + steps, per_sweep, per_sweep * 1e6 / n_markers),
            "s_per_sweep_sample": per_sweep, "marker_build_s": t["build_s"]}


def main():
    a = parse()
    # The contract is ONE JSON line on stdout: keep everything else (NCCL's version banner, library chatter) away from it by
    # pointing fd 1 at stderr for the duration of the run and writing the line to the saved descriptor.
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)

    def emit(line):
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    N, M = a.n_inds, a.n_markers
    workload = "synthetic dense 2-bit N=%d x M=%d, MAF~U(%.2f,%.2f), %.1f%% missing, %d causal h2=%.1f, K=4 S=(0.01,0.001,0.0001)" % (
        N, M, MAF_LO, MAF_HI, 100 * P_MISSING, N_CAUSAL, H2)

    if a.impl == "reference":
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_lib as R
        if not (R.ref_available(fast=True) or R.ref_available()):
            emit({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"})
            return
        y = reference_phenotype(N, M)          # y = X beta + e, as the engine arm
        leg = cpu_reference_leg(N, a.cpu_markers, y, a.steps, a.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": leg["value"],
                "unit": "marker updates/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * leg["s_per_sweep_sample"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": leg["sample"], "arithmetic": "fp64 (Eigen), -O3 -ffast-math as shipped"},
                "cpu_baseline": leg,
                "e2e": {"value": leg["value"], "unit": "marker updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from bayesrrcmd_b200 import Engine

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    # ---- shard individuals over ranks (bayesrrcmd_b200/sharding.py) -------------------------------------------
    from bayesrrcmd_b200.sharding import ShardGroup, shard_bounds
    grp = ShardGroup(rank, world)
    lo, hi = shard_bounds(N, world)[rank]
    n_loc = hi - lo
    eng = Engine(n_loc, M, 4, 1, layout="dense", device=local_rank, block_markers=a.block, n_inds_total=N, max_ctas=int(os.environ.get("BRR_BENCH_MAX_CTAS", 0)),
                 rank=rank, world=world)
    info = eng.info()
    t0 = time.time()
    eng.synth_dense(SEED, MAF_LO, MAF_HI, False, P_MISSING, ind_offset=lo)
    counts = grp.allreduce_counts(eng.counts())
    if world > 1:
        eng.ipc_attach(grp.gather(eng.ipc_export()))
    eng.finalize(counts)
    eng.set_model(S_CVA)
    gen_s = time.time() - t0

    # ---- phenotype: y = X beta + e with N_CAUSAL causal markers, h2 = 0.5, standardised like the reference
    rng = np.random.default_rng(SEED)          # identical stream on every rank
    causal = rng.choice(M, N_CAUSAL, replace=False)
    bcausal = rng.standard_normal(N_CAUSAL) * np.sqrt(H2 / N_CAUSAL)
    noise = rng.standard_normal(N) * np.sqrt(1.0 - H2)
    eng.set_state(eps=noise[lo:hi], beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
    for j, b in zip(causal, bcausal):
        eng.apply_delta(int(j), float(b), 0.0)            # eps += b * x_j
    y = grp.standardise(eng.get_state(beta=False, comp=False, acum=False)[0], N)
    eng.set_state(eps=y)
    y_full = np.concatenate(grp.gather(y))
    allsum = grp.allsum

    # ---- chain state as BayesRBase::init leaves it (reference src/BayesRBase.cpp:39-99) -------------
    K = 4
    pi = np.full((1, K), 0.5 / K); pi[0, 0] = 0.5
    sigmaG = np.array([rng.random()])
    sigmaE = allsum(float((y ** 2).sum())) / N * 0.5
    mu = 0.0

    kernel_ms, e2e_s, nonskipped, blocks = [], [], [], []
    warm = []                                  # (kernel ms, non-skipped markers, dense kernel version) of the warm-up sweeps: the chain's update-heavy start
    kernels, restarts = [], []                 # timed sweeps: dense kernel version the library chose, pipeline restarts (version 5)
    state_hash = [None]
    h2d = d2h = 0
    sampler = None

    # the step's host buffers live in pinned memory (inputs: marker order + (u, z) table; outputs: beta, components, acum)
    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()
    rand_uz = pinned((M, 2), torch.float64)
    order = pinned((M,), torch.int32).view(np.uint32)
    outbuf = {"beta": pinned((M,), torch.float64), "comp": pinned((M,), torch.float64), "acum": pinned((M,), torch.float64)}

    def one_sweep(timed):
        nonlocal mu, sigmaE, sigmaG, pi, h2d, d2h
        s = allsum(eng.eps_add_scalar(mu))                      # updateMu (src/DenseBayesRRmz.cpp:36-42)
        mu = rng.normal(s / N, np.sqrt(sigmaE / N))
        eng.eps_add_scalar(-mu)
        rand_uz[:, 0] = rng.random(M); rand_uz[:, 1] = rng.standard_normal(M)
        order[:] = rng.permutation(M).astype(np.uint32)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, bs, out = eng.sweep(order, rand_uz, pi, sigmaG, sigmaE)     # the plugin call: host buffers in, host out
        _, beta, comp, acum = eng.get_state(eps=False, out=outbuf)
        t1 = time.perf_counter()
        state_hash[0] = (beta.copy(), comp.copy())
        kv, rst = eng.last_sweep_info()
        if not timed:
            warm.append((out.kernel_ms, out.n_nonskipped, kv))
        else:
            kernels.append(kv); restarts.append(rst)
        if timed:
            kernel_ms.append(out.kernel_ms); e2e_s.append(t1 - t0)
            nonskipped.append(out.n_nonskipped); blocks.append(out.n_blocks)
            h2d = order.nbytes + rand_uz.nbytes + pi.nbytes + sigmaG.nbytes
            d2h = v.nbytes + bs.nbytes + 32 + beta.nbytes + comp.nbytes + acum.nbytes
        sq = allsum(out.eps_sqnorm)
        sigmaE = inv_scaled_chisq(rng, V0E + N, (sq + V0E * S02E) / (V0E + N))           # src/BayesRBase.cpp:238-239
        m0 = v[0, 1:].sum()
        sigmaG = np.array([inv_scaled_chisq(rng, V0G + m0, (bs[0] * m0 + V0G * S02G) / (V0G + m0))])   # :243-245
        pi = rng.dirichlet(v[0] + 1.0)[None, :]                                          # :246
        return out

    for _ in range(a.warmup):
        one_sweep(False)
    launches0 = eng.launch_count()
    profiling = os.environ.get("BRR_GPU_PROFILE") == "1"
    if profiling:
        eng.profile_cycles(reset=True)
        eng.profile_cycles5(reset=True)
        eng.profile_cycles6(reset=True)
    if rank == 0:
        sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(a.steps):
        one_sweep(True)
    launches = eng.launch_count() - launches0
    if sampler:
        sampler.stop_flag = True; sampler.join(timeout=2)

    if profiling and rank == 0 and 5 in kernels:      # version-5 kernel: cycles per walked block of one warp per role of CTA 0 -> stderr
        pc = eng.profile_cycles5(reset=True)
        nblk = max(1, pc["blocks"] if pc.get("blocks") else int(np.sum(blocks)))
        sys.stderr.write("v5 phase cycles/block over %d timed sweeps (%d blocks): " % (a.steps, int(np.sum(blocks)))
                         + "  ".join("%s=%d" % (k, pc[k] // max(1, int(np.sum(blocks)))) for k in eng.PROFILE5 if k not in ("blocks", "gram_blocks", "restarts", "applies"))
                         + "  | per block: gram_blocks=%.3f applies=%.3f  restarts/sweep=%.1f\n" % (
                             pc["gram_blocks"] / max(1, int(np.sum(blocks))), pc["applies"] / max(1, int(np.sum(blocks))), float(np.mean(restarts))))
    if profiling and rank == 0 and 6 in kernels:      # version-6 kernel: cycles per block of the sweep, one warp per role of CTA 0 -> stderr
        pc = eng.profile_cycles6(reset=True)
        nblk = max(1, int(np.sum(blocks)))
        cnts = ("blocks", "walks", "applies", "restarts", "gram_blocks", "row_blocks", "c_iters")
        sys.stderr.write("v6 phase cycles/block over %d timed sweeps (%d blocks): " % (a.steps, nblk)
                         + "  ".join("%s=%d" % (k, pc[k] // nblk) for k in eng.PROFILE6 if k not in cnts and not k.startswith("w_"))
                         + " (d_full = d_pre [release of the set, snapshot] + d_fullonly [rows] + wait for stashed rows; d_loop = blocks without Gram rows, d_loopgram = with)"
                         + "  | per walk of CTA 0 (%d walks): " % pc["walks"] + "  ".join("%s=%d" % (k, pc[k] // max(1, pc["walks"])) for k in eng.PROFILE6 if k.startswith("w_"))
                         + "  | per block: applies=%.3f  Gram rows on a dot pass=%.3f  walks with a row=%.3f of CTA 0's  restarts/sweep=%.1f  control-warp passes=%.2f\n" % (
                             pc["applies"] / nblk, pc["row_blocks"] / nblk, pc["gram_blocks"] / max(1, pc["walks"]), float(np.mean(restarts)), pc["c_iters"] / nblk))
    if profiling and rank == 0 and 4 in kernels:      # phase breakdown of the timed sweeps (cycles of CTA 0 per block) -> stderr
        pc = eng.profile_cycles(reset=True)
        nblk = max(1, int(np.sum(blocks)))
        skip = set(eng.PROFILE_COUNTERS) | set(eng.PROFILE_OVERLAPPED) | {"prologue"}
        sys.stderr.write("phase cycles/block over %d timed sweeps (%d blocks): " % (a.steps, nblk)
                         + "  ".join("%s=%d" % (k, pc[k] // nblk) for k in eng.PROFILE_PHASES if k not in skip)
                         + "  | total=%d" % (sum(pc[k] for k in eng.PROFILE_PHASES if k not in skip) // nblk)
                         + "  | overlapped: " + "  ".join("%s=%d" % (k, pc[k] // nblk) for k in eng.PROFILE_OVERLAPPED)
                         + "  | per block: " + "  ".join("%s=%.3f" % (k, pc[k] / nblk) for k in eng.PROFILE_COUNTERS) + "\n")
    ksum = sum(kernel_ms) / 1e3; esum = sum(e2e_s)
    ranks_identical = None
    if world > 1:
        t = torch.tensor([ksum, esum], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ksum, esum = float(t[0]), float(t[1])
        # every rank walks the same chain from the same integer sums: beta / components must be BIT-identical on all ranks
        import hashlib
        h = hashlib.sha256(state_hash[0][0].tobytes() + state_hash[0][1].tobytes()).hexdigest()
        hs = grp.gather(h)
        ranks_identical = all(x == hs[0] for x in hs)
        if not ranks_identical:
            raise SystemExit("bench: beta/components differ between ranks after the timed sweeps: %s" % hs)
    # NCCL baseline of the per-block exchange (north_star: "in-kernel P2P flag-based exchange or an NCCL allreduce baseline, each
    # measured"): the 258 64-bit sums a block reduces across GPUs, as one ncclAllReduce(sum, int64) per block, back to back on one
    # stream -- the exchange ALONE, without any of the sweep's arithmetic.  The in-kernel exchange is inside `ms_per_step`.
    nccl_base = None
    if world > 1:
        buf = torch.zeros(258, dtype=torch.int64, device="cuda")
        for _ in range(50):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        n_ar = 2000
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        ev0.record()
        for _ in range(n_ar):
            dist.all_reduce(buf)
        ev1.record(); torch.cuda.synchronize()
        us = 1e3 * ev0.elapsed_time(ev1) / n_ar
        t = torch.tensor([us], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nblk = int(np.mean(blocks))
        nccl_base = {"what": "ncclAllReduce(sum) of the 258 int64 sums of one block, %d calls back to back on one stream (torch.distributed, NCCL)" % n_ar,
                     "us_per_allreduce": float(t[0]), "blocks_per_sweep": nblk,
                     "exchange_only_s_per_sweep": float(t[0]) * 1e-6 * nblk,
                     "in_kernel_p2p_whole_sweep_s": ksum / a.steps}
    if rank != 0:
        dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    kv_timed = max(set(kernels)) if kernels else 4              # dense kernel version of the timed sweeps (the library chooses per sweep)
    kname = {4: "sweep_dense_kernel", 5: "sweep_dense5_kernel", 6: "sweep_dense6_kernel"}.get(kv_timed, "sweep_dense_kernel")
    # DRAM bytes of one full-size launch: from the committed ncu capture (profiles/ncu_traffic.json), only when that capture
    # was taken on THIS kernel version and problem size (null otherwise: a stale figure explains nothing)
    traffic, traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if world == 1 and N == int(t.get("n_inds", 0)) and M == int(t.get("n_markers", 0)) and t.get("kernel_version") == "v%d" % kv_timed:
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
            traffic_src = "ncu --set full capture of one launch of kernel version %s, committed as profiles/ncu_traffic.json" % t.get("kernel_version")
    except Exception:
        pass
    alg_bytes = ((n_loc + 3) // 4) * M                  # ceil(N_shard/4) genotype bytes per marker update
    achieved = alg_bytes / (ksum / a.steps) / 1e9
    line = {
        "metric": METRIC,
        "value": M * a.steps / ksum, "unit": "marker updates/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * ksum / a.steps, "s_per_sweep": ksum / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "parallelism": "individuals sharded over %d GPU(s)" % world,
                   "block_markers": info["block_markers"], "ctas": info["n_ctas"], "slice_inds": info["slice_inds"],
                   "genotype_bytes_per_gpu": info["genotype_bytes"], "l2": "inputs (%.1f GB/GPU) are larger than L2" % (info["genotype_bytes"] / 1e9),
                   "n_inds": N, "n_markers": M, "arithmetic": ARITHMETIC, "kernel_version": "v%d" % kv_timed,
                   "regime": "timed sweeps = sweeps %d..%d of a chain started from BayesRBase::init's state; %.4f%% of the markers non-skipped "
                             "(beta_old != 0 or beta_new != 0) per timed sweep; the cost of a sweep depends on that fraction" % (
                                 a.warmup, a.warmup + a.steps - 1, 100.0 * float(np.mean(nonskipped)) / M),
                   "chain_start": [{"sweep": i, "ms": float(ms), "nonskipped_fraction": ns / M, "dense_kernel": kv} for i, (ms, ns, kv) in enumerate(warm[:3])],
                   "dense_kernel_timed_sweeps": sorted(set(kernels)), "pipeline_restarts_per_sweep": float(np.mean(restarts)) if restarts else None,
                   "nonskipped_fraction": float(np.mean(nonskipped)) / M, "exchanges_per_sweep": int(np.mean(blocks)),
                   "ranks_bit_identical": ranks_identical, "setup_s": gen_s},
        "e2e": {"value": M * a.steps / esum, "unit": "marker updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "s_per_sweep": esum / a.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kname,
                     "algorithmic_bytes_per_launch": int(alg_bytes)},
        "clocks": sampler.summary() if sampler else None,
    }
    if nccl_base:
        line["exchange"] = {"in_kernel": "integer sums added into count-carrying accumulators of every peer over NVLink inside the sweep kernel",
                            "nccl_allreduce_baseline": nccl_base}
    if not a.no_cpu_baseline:
        try:
            leg = cpu_reference_leg(N, a.cpu_markers, y_full, a.cpu_sweeps)
        except Exception as ex:      # the checker must never take the product's number down with it
            leg = {"error": repr(ex)}
        line["cpu_baseline"] = leg
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
