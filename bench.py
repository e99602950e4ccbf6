#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric: marker-updates/s per Gibbs iteration on the UK-Biobank-shaped
synthetic workload (N=458,000 individuals, M=1,000,000 markers, 1 trait), marker-sharded over N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm (one process per GPU)
  python bench.py --impl reference [--gpus N] --steps K --warmup W    the reference's CPU path (oracle/_ref)

A "step" is one Gibbs iteration (every marker updated once).  One JSON line is printed by rank 0.
Timing: CUDA events inside the engine (device time of each iteration, max over ranks); the genotype
matrix (>= 100 GB at full size) is far larger than L2, so no flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, M, T, G  (BASELINE.json configs)
    "ukb": dict(N=458_000, M=1_000_000, T=1, G=1),          # configs[2]: the config the metric is quoted on
    "c2": dict(N=20_000, M=50_000, T=1, G=1),               # configs[1]
    "c4": dict(N=200_000, M=500_000, T=4, G=1),             # configs[3]
    "c5": dict(N=458_000, M=1_000_000, T=1, G=20),          # configs[4]
    "ukbn": dict(N=458_000, M=50_000, T=1, G=1),            # UKB individuals, C2's marker count (long-chain probes)
    "tiny": dict(N=20_000, M=8_192, T=1, G=1),              # quick functional run
}
MIXTURES = (0.0, 1e-4, 1e-3, 1e-2)                           # example/test.grm


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def full_mask(N):
    mask4 = np.full((N + 3) // 4, 0x0F, dtype=np.uint8)
    if N % 4:
        mask4[-1] = (1 << (N % 4)) - 1
    return mask4


def causal_effects(M, T, causal_frac, h2, seed):
    """data_sim.R:17-33: ncausal = causal_frac * M markers (5,000 of 20,000 there), effects ~ N(0, h2 / ncausal)."""
    ncausal = max(1, int(round(causal_frac * M)))
    beta = np.zeros((T, M))
    for t in range(T):
        rng = np.random.default_rng(seed + 1000 * t)
        idx = rng.choice(M, ncausal, replace=False)
        beta[t, idx] = rng.normal(0.0, np.sqrt(h2 / ncausal), size=ncausal)
    return beta, ncausal


def phenotype_from_engine(e, N, M, T, causal_frac, h2, seed):
    """y = scale(X) b + e (example/data_sim.R:17-41) with the genetic values taken ON THE DEVICE over every
    marker of every shard (gmrm_genetic_values: one pass over the genotypes + one all-reduce).  Every rank draws the same
    effects and noise from the same seeds, so all ranks end up with the same phenotype."""
    mask4 = full_mask(N)
    for t in range(T):                                   # marker means / scales under the all-observed mask
        e.set_phenotype(t, np.zeros(N), mask4, N)
    e.compute_marker_stats()
    beta, ncausal = causal_effects(M, T, causal_frac, h2, seed)
    ys = []
    for t in range(T):
        g = e.genetic_values(t, beta[t, e.marker_begin:e.marker_begin + e.marker_count])
        rng = np.random.default_rng(seed + 1000 * t + 1)
        ys.append(g + rng.normal(0.0, np.sqrt(max(1e-6, 1.0 - g.var())), N))
    return np.stack(ys), ncausal


def standardise(y, na):
    """Phenotype::read_file's centring/scaling (phenotype.cpp:647-667) for an in-memory phenotype."""
    obs = ~na
    c = np.where(obs, y - y[obs].mean(), 0.0)
    c *= np.sqrt((obs.sum() - 1) / (c ** 2).sum())
    mask4 = np.zeros((y.size + 3) // 4, dtype=np.uint8)
    idx = np.nonzero(obs)[0]
    np.bitwise_or.at(mask4, idx // 4, (1 << (idx % 4)).astype(np.uint8))
    return c, mask4, int(obs.sum())


def ncu_traffic_per_launch(vranks_per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel launch, from the newest committed `ncu --set full`
    capture of this same command (profiles/r*_step_kernel_traffic.json); None if no capture matches."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_kernel_traffic.json")), reverse=True):
        d = json.load(open(p))
        if d.get("vranks_per_gpu") == vranks_per_gpu:
            return d.get("dram_bytes_per_launch")
    return None


def multi_gpu_parity_check(rank, world, local):
    """Before anything is timed on N > 1 GPUs: the small marker-sharded chain of tests/mgpu_check.py (2 traits, NAs,
    missing genotypes, 2 groups) at sync_rate 1 and 3 against the oracle with the same total number of virtual ranks;
    at sync_rate 1 every rank's residuals must be BIT-identical.  The oracle is the checker here, nothing timed."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mgpu_check
    out = {"ok": True, "cases": []}
    for sync_rate in (1, 3):
        try:
            info = mgpu_check.chain_check(rank, world, local, sync_rate)
            info.pop("inputs", None)
            info["bit_identical_replicas"] = sync_rate == 1 and os.environ.get("GMRM_EXCHANGE") != "delta"
            out["cases"].append(info)
        except Exception as ex:                      # noqa: BLE001 -- the bench line must say what failed
            out["ok"] = False
            out["cases"].append({"sync_rate": sync_rate, "error": f"{type(ex).__name__}: {ex}"[:400]})
    return out


def setup_probes(api, e, N, M, T, local, peak):
    """One-off costs next to the per-iteration figures, rank 0: marker statistics over the resident
    shard, the association pass on a slice, ingestion from host memory.  Host wall clock around synchronous C-ABI calls."""
    mbytes = (N + 3) // 4
    out = {}
    t = time.perf_counter()
    e.compute_marker_stats()
    dt = time.perf_counter() - t
    b = e.marker_count * mbytes * T
    out["marker_stats"] = {"seconds": dt, "markers": e.marker_count, "traits": T, "alg_bytes": b, "gbs": b / dt / 1e9,
                           "frac_of_hbm_peak": b / dt / 1e9 / peak, "ref": "phenotype.cpp:466-556"}
    # association pass (--predict, bayes.cpp:14-284) and ingestion on a separate 32,768-marker engine
    Ms = min(M, 32768)
    eu = api.Engine(N=N, Mt=Ms, vranks=8, device=local)
    eu.generate_bed(seed=2, missing_rate=0.002)
    eu.finalize_bed()
    rng = np.random.default_rng(11)
    y = rng.normal(size=N)
    y = (y - y.mean()) / y.std()
    eu.set_phenotype(0, y, full_mask(N), N)
    eu.compute_marker_stats()
    beta = rng.normal(0, 0.01, size=Ms) * (rng.random(Ms) < 0.05)
    eu.predict(0, y, beta)                             # warm-up: allocations, first launches
    t = time.perf_counter()
    eu.predict(0, y, beta)
    dt = time.perf_counter() - t
    out["predict"] = {"seconds": dt, "markers": Ms, "blocks": 8, "passes": 3, "alg_bytes": 3 * Ms * mbytes,
                      "gbs": 3 * Ms * mbytes / dt / 1e9, "frac_of_hbm_peak": 3 * Ms * mbytes / dt / 1e9 / peak,
                      "ref": "bayes.cpp:87-214", "note": "host wall clock: uploads of y / beta and read-back of 4 x M doubles included"}
    up_host = eu.download_bed()
    up_pinned = api.host_array(up_host.shape)
    up_pinned[...] = up_host
    rates = []
    for src in (up_host, up_pinned):                   # pageable numpy buffer, then pinned (gmrm_host_alloc)
        t_up = time.perf_counter()
        eu.upload_bed(src)
        eu.finalize_bed()
        rates.append(round(src.nbytes / (time.perf_counter() - t_up) / 1e9, 2))
    full = M * mbytes
    out["upload_bed"] = {"sample_bytes": int(up_host.nbytes), "pageable_gbs": rates[0], "pinned_gbs": rates[1],
                         "full_matrix_bytes": full, "full_matrix_s_at_pinned_rate": round(full / (rates[1] * 1e9), 2),
                         "ref": "bayes.cpp:867-900",
                         "note": "one-time ingestion of the whole matrix, not part of value / e2e (an iteration has no host inputs); "
                                 "profiles/r2_cli_ingest.json holds a run of the executable from a real 45.8 GB .bed file (5.9 GB/s end to end: file read into pinned memory + upload)"}
    eu.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gmrm_b200 import api

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parity = None
    if world > 1 and not args.no_parity_check:
        parity = multi_gpu_parity_check(rank, world, local)
    w = WORKLOADS[args.workload]
    N, M, T, G = w["N"], args.markers or w["M"], w["T"], w["G"]
    K = len(MIXTURES)
    R = args.vranks_per_gpu * world
    e = api.Engine(N=N, Mt=M, T=T, G=G, K=K, vranks=R, world_size=world, world_rank=rank, sync_rate=args.sync_rate,
                   seed=171014, device=local)
    if world > 1:
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        e.comm_init(uid[0])
    t0 = time.time()
    e.generate_bed(seed=1, maf_lo=0.05, maf_hi=0.5, missing_rate=0.0)
    e.finalize_bed()
    if world > 1 and args.sync_rate == 1:      # list exchange: the shards read each other's columns over NVLink

        def gather(x):
            out = [None] * world
            dist.all_gather_object(out, x)
            return out
        e.exchange_buffers(gather)
    # phenotype: y = scale(X) b + e, causal_frac * M causal markers, h2 = 0.5; genetic values on the device
    y, ncausal = phenotype_from_engine(e, N, M, T, args.causal_frac, 0.5, seed=171014)
    rng = np.random.default_rng(3)
    for t in range(T):
        na = rng.random(N) < (0.01 if T > 1 else 0.0)
        c, mask4, nonas = standardise(y[t], na)
        e.set_phenotype(t, c, mask4, nonas)
    groups = np.random.default_rng(4).integers(0, G, size=M).astype(np.int32) if G > 1 else np.zeros(M, dtype=np.int32)
    cva = np.stack([np.array(MIXTURES) * (1.0 + g) for g in range(G)])
    e.set_groups(groups, cva)
    e.compute_marker_stats()
    e.init_chain(None)
    setup_s = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumreduce(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    it = 0
    pub_warm, pub_timed = [], []
    for _ in range(args.warmup):
        it += 1
        e.run_iteration(it)
        pub_warm.append(e.timing()["published"])
    # ---- timed: K iterations, device time (CUDA events in the engine around the whole iteration), max over ranks
    e.set_timing_detail(0)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms, dot_ms, dev1_ms, launches = [], [], [], 0
    for _ in range(args.steps):
        it += 1
        e.run_iteration(it)
        tm = e.timing()
        dev_ms.append(tm["iteration_ms"])
        launches += tm["launches"]; pub_timed.append(tm["published"])
    barrier()
    ms_per_step = maxreduce(sum(dev_ms) / len(dev_ms))
    # ---- e2e: the call a user makes per iteration -- run it, then read the iteration's outputs back to the
    # host (what the reference writes to .bet/.cpn/.csv, bayes.cpp:659-669); host wall clock, max over ranks.
    # Everything is read through the staged path: device snapshot, pinned D2H on a second stream, fetched one iteration later.
    # one untimed pass of the same call sequence first: the staging buffers (pinned host memory, device snapshots, copy
    # stream) are allocated on first use -- set-up, like the chain's own warm-up iterations
    it += 1
    e.run_iteration_async(it)
    e.wait_iteration()
    e.stage_outputs()
    for t in range(T):
        e.fetch_outputs(t)
    e.fetch_state()
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    staged = False
    for _ in range(args.steps):
        it += 1
        e.run_iteration_async(it)                    # enqueue; the host reads the previous iteration's outputs meanwhile
        if staged:                                   # outputs of the previous iteration: their copy ran ahead of this one
            for t in range(T):
                bb, cc = e.fetch_outputs(t)
                d2h += bb.nbytes + cc.nbytes
            d2h += sum(v.nbytes for v in e.fetch_state().values())
        e.wait_iteration()
        e.stage_outputs()
        staged = True
    for t in range(T):
        bb, cc = e.fetch_outputs(t)
        d2h += bb.nbytes + cc.nbytes
    st = e.fetch_state()
    d2h += sum(v.nbytes for v in st.values())
    barrier()
    e2e_s = maxreduce((time.perf_counter() - t1) / args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # (the diagnostic passes below come after both timed regions, so that the two see the chain at neighbouring iterations)
    # K further iterations with 2 CUDA events per step around the step kernel: the live launch duration behind `roofline`
    # (events between the launches switch the programmatic overlap of consecutive kernels off, hence not the K above)
    e.set_timing_detail(1)
    for _ in range(args.steps):
        it += 1
        e.run_iteration(it)
        tm = e.timing()
        dot_ms.append(tm["dot_kernel_ms"]); dev1_ms.append(tm["iteration_ms"])
    barrier()
    # diagnostic pass (not part of the K timed steps): one iteration with every phase bracketed by events and the
    # residual update of every step as its own launch, so that dot-only and update-only times are seen
    e.set_timing_detail(2)
    it += 1
    e.run_iteration(it)
    tm2 = e.timing()
    e.set_timing_detail(0)
    pub_warm = [sumreduce(x) for x in pub_warm]
    pub_timed = [sumreduce(x) for x in pub_timed]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        mbytes = (N + 3) // 4
        steps_per_it = e.Mm
        # algorithmic bytes of one step-kernel launch: V columns of ceil(N/4) packed bytes + the residuals once
        alg_bytes_launch = args.vranks_per_gpu * mbytes + 8 * N * T
        avg_dot_ms = (sum(dot_ms) / len(dot_ms)) / steps_per_it
        achieved = alg_bytes_launch / (avg_dot_ms * 1e-3) / 1e9
        dot_only_ms = tm2["dot_kernel_ms"] / steps_per_it
        # whole-iteration algorithmic traffic: genotype stream + per-step residual pass + marker scalars
        it_bytes = M * mbytes + steps_per_it * 2 * 8 * N * T * world + 36 * M * T
        out = {
            "metric": "marker-updates/sec per Gibbs iter (UKB shape)", "value": M * T / (ms_per_step * 1e-3),
            "unit": "marker-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": f"synthetic (Binomial(2,p) genotypes, p~U(0.05,0.5), generated on device; phenotype y = scale(X) b + e "
                                    f"with {ncausal} causal markers, h2 = 0.5, genetic values computed on device)",
            "config": {"workload": f"{args.workload}: N={N} M={M} T={T} G={G} K={K}", "vranks_per_gpu": args.vranks_per_gpu,
                       "vranks_total": R, "sync_rate": args.sync_rate, "marker_steps_per_iter": steps_per_it,
                       "exchange": None if world == 1 else (os.environ.get("GMRM_EXCHANGE") or "xdelta") if args.sync_rate == 1 else "delta",
                       "causal_markers": ncausal, "h2": 0.5,
                       "layout": f"base-3 quads (1 byte = 4 genotypes), {e.tiles} CTAs, {e.column_stride} B/column",
                       "l2": "inputs >> L2 (no flush needed)",
                       "setup_s": round(setup_s, 1), "hbm_gbs_iter": it_bytes / (ms_per_step * 1e-3) / 1e9,
                       "hbm_frac_iter": it_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world)},
            "roofline": {"bound": "hbm", "kernel": "step_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_per_launch(args.vranks_per_gpu), "peak_source": peak_src,
                         "dot_share_of_step": (sum(dot_ms) / len(dot_ms)) / (sum(dev1_ms) / len(dev1_ms)),
                         "alg_bytes_per_launch": alg_bytes_launch, "avg_launch_ms": avg_dot_ms,
                         "frac_dot_only": alg_bytes_launch / (dot_only_ms * 1e-3) / 1e9 / peak if dot_only_ms > 0 else None,
                         "per_step_us": {"dot": 1e3 * avg_dot_ms, "step": 1e3 * ms_per_step / steps_per_it,
                                         "dot_only": 1e3 * dot_only_ms, "update": 1e3 * tm2["update_kernel_ms"] / steps_per_it,
                                         "sample": 1e3 * tm2["sample_kernel_ms"] / steps_per_it,
                                         "exchange": 1e3 * tm2["exchange_ms"] / steps_per_it,
                                         "allreduce": 1e3 * tm2["allreduce_ms"] / steps_per_it,
                                         "published_per_step_diag": tm2["published"] / steps_per_it,
                                         "step_with_events": 1e3 * (sum(dev1_ms) / len(dev1_ms)) / steps_per_it,
                                         "note": "step: the K timed iterations (no events inside); dot (step kernel: pending residual updates + table "
                                                 "build + dot products): K further iterations with 2 events per step; "
                                                 "dot_only/update/sample/exchange: one extra iteration with 6 events per step and the "
                                                 f"residual update as its own launch ({1e3 * tm2['iteration_ms'] / steps_per_it:.1f} us per step in that pass)"}},
            "e2e": {"value": M * T / e2e_s, "unit": "marker-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": d2h // args.steps,
                    "note": "per-iteration calls through the C ABI (gmrm_run_iteration_async / gmrm_wait_iteration) + read-back of "
                            "betas/components/state (staged: device snapshot, pinned D2H on a second stream, fetched while the next iteration runs) to host "
                            "buffers (what the reference writes to .bet/.cpn/.csv); an iteration has no host inputs: genotypes and "
                            "phenotypes are uploaded once per run (setup.upload_bed is that path's measured rate)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "chain": {"sigmaE": float(st["sigmae"][0]), "sigmaG_sum": float(st["sigmag"][0].sum()),
                      "published_per_iter": sum(pub_timed) / len(pub_timed), "published_warmup": pub_warm, "published_timed": pub_timed,
                      "published_frac": sum(pub_timed) / len(pub_timed) / (M * T)},
        }
        if parity is not None:
            out["parity_check"] = parity
        if world == 1 and not args.no_setup_probes:
            out["setup"] = setup_probes(api, e, N, M, T, local, peak)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(e, N, M, T, G, args)
        print(json.dumps(out), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# Reference legs (the checker / baseline side): nothing below touches the product library.
_DOSAGE_OF_CODE = np.array([2, 3, 1, 0], dtype=np.uint8)       # PLINK code -> dosage, 3 = missing (lut/mk_lut.cpp:25-32)


def host_bed(N, Ms, seed=1, maf_lo=0.05, maf_hi=0.5):
    """Packed PLINK bytes of Ms synthetic markers (Binomial(2, p), p ~ U(maf_lo, maf_hi)), numpy only, marker chunks."""
    from gmrm_b200 import synth
    rng = np.random.default_rng(seed)
    p = rng.uniform(maf_lo, maf_hi, size=Ms)
    out = np.empty((Ms, (N + 3) // 4), dtype=np.uint8)
    for j0 in range(0, Ms, 256):
        pj = p[j0:j0 + 256, None].astype(np.float32)
        d = (rng.random((pj.shape[0], N), dtype=np.float32) < pj).astype(np.uint8)
        d += rng.random((pj.shape[0], N), dtype=np.float32) < pj
        out[j0:j0 + 256] = synth.pack_bed(d)
    return out


def host_phenotypes(bed, N, T, causal_frac, h2, seed):
    """y = scale(X) b + e from the slice's own columns (data_sim.R:17-41): same causal fraction and h2 as the GPU arm's."""
    Ms = bed.shape[0]
    beta, ncausal = causal_effects(Ms, T, causal_frac, h2, seed)
    ys = []
    for t in range(T):
        g = np.zeros(N)
        for j in np.nonzero(beta[t])[0]:
            codes = np.stack([(bed[j] >> (2 * k)) & 3 for k in range(4)], axis=1).reshape(-1)[:N]
            d = _DOSAGE_OF_CODE[codes].astype(np.float64)
            obs = d < 3
            z = np.where(obs, d - d[obs].mean(), 0.0)
            sd = z[obs].std(ddof=1)
            g += beta[t, j] * z / (sd if sd > 0 else 1.0)
        rng = np.random.default_rng(seed + 1000 * t + 1)
        ys.append(g + rng.normal(0.0, np.sqrt(max(1e-6, 1.0 - g.var())), N))
    return np.stack(ys), ncausal


def write_sample_files(tmp, bed, y, N, T, G, seed=5):
    Ms = bed.shape[0]
    with open(os.path.join(tmp, "s.bed"), "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01])); f.write(bed.tobytes())
    open(os.path.join(tmp, "s.dim"), "w").write(f"{N} {Ms}\n")
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, G, size=Ms) if G > 1 else np.zeros(Ms, dtype=int)
    open(os.path.join(tmp, "s.gri"), "w").write("".join(f"{j} {int(g)}\n" for j, g in enumerate(grp)))
    open(os.path.join(tmp, "s.grm"), "w").write("".join(" ".join(f"{v * (1.0 + g):.5f}" for v in MIXTURES) + "\n" for g in range(G)))
    phens = []
    ids = np.arange(1, N + 1)
    na_rng = np.random.default_rng(3)
    for t in range(T):
        p = os.path.join(tmp, f"s_t{t}.phen")
        if T > 1:                                    # 1 % NAs per trait, literal NA (test1_nas.phen style)
            na = na_rng.random(N) < 0.01
            with open(p, "w") as f:
                f.write("".join(f"{i} {i} {'NA' if m else format(v, '.10f')}\n" for i, v, m in zip(ids, y[t], na)))
        else:
            np.savetxt(p, np.column_stack([ids, ids, y[t]]), fmt=["%d", "%d", "%.10f"])
        phens.append(p)
    return phens


def time_reference(bed, y, N, T, G, iterations, configs):
    """oracle/_ref/gmrm_ref (the unmodified reference; MPI shim: ranks are threads of one process, each with its own
    OpenMP team) on a marker slice, for every (ranks, threads) in `configs`; returns {(ranks, threads): seconds per
    iteration from its own RESULT lines (bayes.cpp:655), iteration 1 dropped}."""
    from oracle import oracle_py as O
    if not O.have_reference():
        return None
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        phens = write_sample_files(tmp, bed, y, N, T, G)
        for (ranks, threads) in configs:
            out = O.run_reference(tmp, os.path.join(tmp, "s.bed"), os.path.join(tmp, "s.dim"), phens, os.path.join(tmp, "s.gri"),
                                  os.path.join(tmp, "s.grm"), os.path.join(tmp, f"out_{ranks}_{threads}"), iterations=iterations,
                                  seed=171014, nranks=ranks, threads=threads, timeout=3000)
            times = [float(l.split("total proc time =")[1].split("sec")[0]) for l in out.splitlines() if "total proc time" in l]
            res[(ranks, threads)] = times[1:] if len(times) > 1 else times
    return res


def reference_configs(cores):
    """(MPI ranks, OpenMP threads per rank) splits of the host cores the reference is timed with; the best one is reported."""
    cfgs = [(1, cores)]
    for r in (2, 4, 8):
        if cores >= 2 * r:
            cfgs.append((r, cores // r))
    return cfgs


def cpu_baseline(e, N, M, T, G, args):
    cores = os.cpu_count() or 1
    Ms = min(e.marker_count, args.cpu_markers)
    bed = e.download_bed(e.marker_begin, Ms)
    y, ncausal = host_phenotypes(bed, N, T, args.causal_frac, 0.5, seed=171014)
    res = time_reference(bed, y, N, T, G, 4, reference_configs(cores))
    if not res:
        return {"value": None, "unit": "marker-updates/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"}
    med = {k: statistics.median(v) for k, v in res.items() if v}
    (ranks, threads), s = min(med.items(), key=lambda kv: kv[1])
    return {"value": Ms * T / s, "unit": "marker-updates/s", "cores": ranks * threads, "kind": "reference",
            "sample": f"first {Ms} markers of the same matrix (N={N}), phenotype simulated from that slice ({ncausal} causal markers, h2 0.5), "
                      f"reference binary oracle/_ref/gmrm_ref, best of {sorted(med)} (ranks, threads): {ranks} x {threads}, "
                      f"median of iterations 2-4 = {s:.3f} s",
            "all": {f"{r}x{t}": Ms * T / v for (r, t), v in med.items()}}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the host cores, same workload shape,
    each step a bounded marker slice (per-marker cost does not depend on Mt).  This process never loads
    the product library: genotypes and phenotype are made with numpy."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    N, M, T, G = w["N"], args.markers or w["M"], w["T"], w["G"]
    cores = os.cpu_count() or 1
    Ms = min(M, args.cpu_markers)
    bed = host_bed(N, Ms, seed=1)
    y, ncausal = host_phenotypes(bed, N, T, args.causal_frac, 0.5, seed=171014)
    res = time_reference(bed, y, N, T, G, args.steps + args.warmup, reference_configs(cores))
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/gmrm_ref was not built"}), flush=True)
        return
    mean = {k: sum(v[-args.steps:]) / len(v[-args.steps:]) for k, v in res.items() if v}
    (ranks, threads), s = min(mean.items(), key=lambda kv: kv[1])
    v = Ms * T / s
    out = {"impl": "reference", "metric": "marker-updates/sec per Gibbs iter (UKB shape)", "value": v, "unit": "marker-updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s * 1e3 * (M / Ms),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": f"synthetic (numpy Binomial(2,p) genotypes; phenotype y = scale(X) b + e, {ncausal} causal markers of the slice, h2 0.5)",
           "config": {"workload": f"{args.workload}: N={N} M={M} T={T} G={G} K=4", "sample_markers": Ms, "causal_markers": ncausal,
                      "ranks_x_threads": f"{ranks}x{threads}", "tried": {f"{r}x{t}": Ms * T / x for (r, t), x in mean.items()}},
           "cpu_baseline": {"value": v, "unit": "marker-updates/s", "cores": ranks * threads, "kind": "reference",
                            "sample": f"{Ms}-marker slice, N={N}, unmodified reference sources, {ranks} shim ranks x {threads} OpenMP threads "
                                      f"(best of {sorted(mean)})"},
           "e2e": {"value": v, "unit": "marker-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ukb", choices=sorted(WORKLOADS))
    ap.add_argument("--markers", type=int, default=0, help="override M (debug)")
    ap.add_argument("--vranks-per-gpu", type=int, default=2048)
    ap.add_argument("--sync-rate", type=int, default=1)
    ap.add_argument("--cpu-markers", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-setup-probes", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--causal-frac", type=float, default=0.25, help="causal markers / M (data_sim.R: 5,000 of 20,000)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
