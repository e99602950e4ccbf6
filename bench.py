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
    "tiny": dict(N=20_000, M=8_192, T=1, G=1),              # quick functional run
}
MIXTURES = (0.0, 1e-4, 1e-3, 1e-2)                           # example/test.grm


def ncu_traffic_per_launch(vranks_per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel launch, from the committed `ncu --set full`
    capture of this same command (profiles/r1_step_kernel_traffic.json); None if no capture matches."""
    p = os.path.join(ROOT, "profiles", "r1_step_kernel_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d.get("dram_bytes_per_launch") if d.get("vranks_per_gpu") == vranks_per_gpu else None


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


def phenotype_from_engine(e, N, T, seed):
    """y = X b + e on ~0.5% causal markers of the shard (data_sim.R recipe, reduced to what one host
    can afford at this size: effects are drawn for a marker sample decoded back through the C ABI)."""
    rng = np.random.default_rng(seed)
    ncausal = min(e.marker_count, 200)
    ids = rng.choice(e.marker_count, ncausal, replace=False)
    ys = []
    for t in range(T):
        g = np.zeros(N)
        beta = rng.normal(0.0, np.sqrt(0.5 / ncausal), ncausal)
        for b, j in zip(beta, ids):
            a, nm = e.decode_marker(int(j))
            x = a - a[nm > 0].mean()
            sd = x[nm > 0].std()
            g += b * np.where(nm > 0, x / (sd if sd > 0 else 1.0), 0.0)
        ys.append(g + rng.normal(0.0, np.sqrt(max(1e-6, 1.0 - g.var())), N))
    return np.stack(ys)


def standardise(y, na):
    """Phenotype::read_file's centring/scaling (phenotype.cpp:647-667) for an in-memory phenotype."""
    obs = ~na
    c = np.where(obs, y - y[obs].mean(), 0.0)
    c *= np.sqrt((obs.sum() - 1) / (c ** 2).sum())
    mask4 = np.zeros((y.size + 3) // 4, dtype=np.uint8)
    idx = np.nonzero(obs)[0]
    np.bitwise_or.at(mask4, idx // 4, (1 << (idx % 4)).astype(np.uint8))
    return c, mask4, int(obs.sum())


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
    # phenotype: simulated on rank 0 from its shard, shared with the other ranks (epsilon is replicated)
    obj = [None]
    if rank == 0:
        y = phenotype_from_engine(e, N, T, seed=171014)
        obj = [y]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    y = obj[0]
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
    # host -> HBM ingestion rate of gmrm_upload_bed (numpy host buffer: copy + transcode + missing lists) on a separate
    # 8,192-marker engine (0.94 GB, 4 staging chunks): the one-time cost that the per-iteration e2e figure does not contain
    upload_gbs = None
    if rank == 0:
        eu = api.Engine(N=N, Mt=8192, vranks=1, device=local)
        eu.generate_bed(seed=2)
        up_host = eu.download_bed()
        up_pinned = api.host_array(up_host.shape)
        up_pinned[...] = up_host
        rates = []
        for src in (up_host, up_pinned):             # pageable numpy buffer, then pinned (gmrm_host_alloc)
            t_up = time.perf_counter()
            eu.upload_bed(src)
            eu.finalize_bed()
            rates.append(round(src.nbytes / (time.perf_counter() - t_up) / 1e9, 2))
        upload_gbs = {"pageable": rates[0], "pinned": rates[1]}
        eu.close()
        del up_host, up_pinned

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

    it = 0
    for _ in range(args.warmup):
        it += 1
        e.run_iteration(it)
    # ---- timed: K iterations, device time (CUDA events in the engine), max over ranks
    e.set_timing_detail(1)      # 2 CUDA events per step around the step kernel (the full 6-event split costs ~6 % and is taken below)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms, dot_ms, smp_ms, upd_ms, xch_ms, ar_ms, launches, published = [], [], [], [], [], [], 0, 0
    for _ in range(args.steps):
        it += 1
        e.run_iteration(it)
        tm = e.timing()
        dev_ms.append(tm["iteration_ms"]); dot_ms.append(tm["dot_kernel_ms"])
        launches += tm["launches"]; published += tm["published"]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # diagnostic pass (not part of the K timed steps): one iteration with every phase bracketed by events
    e.set_timing_detail(2)
    it += 1
    e.run_iteration(it)
    tm2 = e.timing()
    smp_ms, upd_ms, xch_ms, ar_ms = [tm2["sample_kernel_ms"]], [tm2["update_kernel_ms"]], [tm2["exchange_ms"]], [tm2["allreduce_ms"]]
    dev2_ms = tm2["iteration_ms"]
    e.set_timing_detail(0)
    ms_per_step = maxreduce(sum(dev_ms) / len(dev_ms))
    # ---- e2e: the call a user makes per iteration -- run it, then read the iteration's outputs back to the
    # host (what the reference writes to .bet/.cpn/.csv, bayes.cpp:659-669); host wall clock, max over ranks
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    staged = False
    for _ in range(args.steps):
        it += 1
        e.run_iteration(it)
        if staged:                                   # outputs of the previous iteration: their copy ran under this one
            for t in range(T):
                bb, cc = e.fetch_outputs(t)
                d2h += bb.nbytes + cc.nbytes
        st = e.state()
        d2h += sum(v.nbytes for v in st.values())
        e.stage_outputs()
        staged = True
    for t in range(T):
        bb, cc = e.fetch_outputs(t)
        d2h += bb.nbytes + cc.nbytes
    barrier()
    e2e_s = maxreduce((time.perf_counter() - t1) / args.steps)
    st = e.state()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        mbytes = (N + 3) // 4
        steps_per_it = e.Mm
        # algorithmic bytes of one dot-kernel launch: V columns of ceil(N/4) packed bytes + the residuals once
        alg_bytes_launch = args.vranks_per_gpu * mbytes + 8 * N * T
        avg_dot_ms = (sum(dot_ms) / len(dot_ms)) / steps_per_it
        achieved = alg_bytes_launch / (avg_dot_ms * 1e-3) / 1e9
        # whole-iteration algorithmic traffic: genotype stream + per-step residual pass + marker scalars
        it_bytes = M * mbytes + steps_per_it * 2 * 8 * N * T * world + 36 * M * T
        out = {
            "metric": "marker-updates/sec per Gibbs iter (UKB shape)", "value": M * T / (ms_per_step * 1e-3),
            "unit": "marker-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (Binomial(2,p) genotypes, p~U(0.05,0.5), generated on device; simulated phenotype)",
            "config": {"workload": f"{args.workload}: N={N} M={M} T={T} G={G} K={K}", "vranks_per_gpu": args.vranks_per_gpu,
                       "vranks_total": R, "sync_rate": args.sync_rate, "marker_steps_per_iter": steps_per_it,
                       "layout": f"base-3 quads (1 byte = 4 genotypes), {e.tiles} CTAs, {e.column_stride} B/column",
                       "l2": "inputs >> L2 (no flush needed)",
                       "setup_s": round(setup_s, 1), "upload_bed_gbs": upload_gbs, "hbm_gbs_iter": it_bytes / (ms_per_step * 1e-3) / 1e9,
                       "hbm_frac_iter": it_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world)},
            "roofline": {"bound": "hbm", "kernel": "step_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_per_launch(args.vranks_per_gpu), "peak_source": peak_src,
                         "dot_share_of_step": (sum(dot_ms) / len(dot_ms)) / (sum(dev_ms) / len(dev_ms)),
                         "alg_bytes_per_launch": alg_bytes_launch, "avg_launch_ms": avg_dot_ms,
                         "per_step_us": {"dot": 1e3 * sum(dot_ms) / len(dot_ms) / steps_per_it, "sample": 1e3 * sum(smp_ms) / len(smp_ms) / steps_per_it,
                                         "update": 1e3 * sum(upd_ms) / len(upd_ms) / steps_per_it,
                                         "exchange": 1e3 * sum(xch_ms) / len(xch_ms) / steps_per_it,
                                         "allreduce": 1e3 * sum(ar_ms) / len(ar_ms) / steps_per_it, "step": 1e3 * ms_per_step / steps_per_it,
                                         "note": "dot and step: the K timed iterations; sample/update/exchange: one extra iteration with 6 events "
                                                 f"per step ({1e3 * dev2_ms / steps_per_it:.1f} us per step in that pass)"}},
            "e2e": {"value": M * T / e2e_s, "unit": "marker-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": d2h // args.steps,
                    "note": "per-iteration call through the C ABI (gmrm_run_iteration) + read-back of betas/components (staged: device "
                            "snapshot, pinned D2H on a second stream, fetched one iteration later) and state to host "
                            "buffers (what the reference writes to .bet/.cpn/.csv); an iteration has no host inputs: genotypes and "
                            "phenotypes are uploaded once per run (config.upload_bed_gbs is that path's measured rate)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "chain": {"sigmaE": float(st["sigmae"][0]), "sigmaG_sum": float(st["sigmag"][0].sum()),
                      "published_per_iter": published / args.steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(e, N, T, G, args)
        print(json.dumps(out), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


def write_sample_files(tmp, bed, N, T, G, seed=5):
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
    for t in range(T):
        y = rng.normal(size=N)
        p = os.path.join(tmp, f"s_t{t}.phen")
        np.savetxt(p, np.column_stack([ids, ids, y]), fmt=["%d", "%d", "%.10f"])
        phens.append(p)
    return phens


def time_reference(bed, N, T, G, iterations, threads):
    """oracle/_ref/gmrm_ref (the unmodified reference, MPI shim with 1 rank, OpenMP on `threads` cores) on a
    marker slice; returns seconds per iteration from its own RESULT lines (bayes.cpp:655), iteration 1 dropped."""
    from oracle import oracle_py as O
    if not O.have_reference():
        return None
    with tempfile.TemporaryDirectory() as tmp:
        phens = write_sample_files(tmp, bed, N, T, G)
        out = O.run_reference(tmp, os.path.join(tmp, "s.bed"), os.path.join(tmp, "s.dim"), phens, os.path.join(tmp, "s.gri"),
                              os.path.join(tmp, "s.grm"), os.path.join(tmp, "out"), iterations=iterations, seed=171014, nranks=1,
                              threads=threads, timeout=3000)
    times = [float(l.split("total proc time =")[1].split("sec")[0]) for l in out.splitlines() if "total proc time" in l]
    return times[1:] if len(times) > 1 else times


def cpu_baseline(e, N, T, G, args):
    cores = os.cpu_count() or 1
    Ms = min(e.marker_count, args.cpu_markers)
    bed = e.download_bed(e.marker_begin, Ms)
    times = time_reference(bed, N, T, G, 4, cores)
    if not times:
        return {"value": None, "unit": "marker-updates/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"}
    s = statistics.median(times)
    return {"value": Ms * T / s, "unit": "marker-updates/s", "cores": cores, "kind": "reference",
            "sample": f"first {Ms} markers of the same matrix (N={N}), reference binary oracle/_ref/gmrm_ref, 1 rank x {cores} "
                      f"OpenMP threads, median of iterations 2-4 = {s:.3f} s"}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the host cores, same workload shape,
    each step a bounded marker slice (per-marker cost does not depend on Mt)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    N, M, T, G = w["N"], args.markers or w["M"], w["T"], w["G"]
    cores = os.cpu_count() or 1
    Ms = min(M, args.cpu_markers)
    from gmrm_b200 import synth
    try:
        from gmrm_b200 import api
        e = api.Engine(N=N, Mt=Ms, T=1, G=1, K=4, vranks=1)
        e.generate_bed(seed=1)
        bed = e.download_bed()
        e.close()
    except Exception:
        bed = synth.pack_bed(synth.make_genotypes(N, Ms, seed=1))
    times = time_reference(bed, N, T, G, args.steps + args.warmup, cores)
    if not times:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/gmrm_ref was not built"}), flush=True)
        return
    times = times[-args.steps:]
    s = sum(times) / len(times)
    v = Ms * T / s
    out = {"impl": "reference", "metric": "marker-updates/sec per Gibbs iter (UKB shape)", "value": v, "unit": "marker-updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s * 1e3 * (M / Ms),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.workload}: N={N} M={M} T={T} G={G} K=4", "sample_markers": Ms},
           "cpu_baseline": {"value": v, "unit": "marker-updates/s", "cores": cores, "kind": "reference",
                            "sample": f"{Ms}-marker slice, N={N}, unmodified reference sources, 1 rank x {cores} OpenMP threads"},
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
