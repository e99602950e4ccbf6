"""Multi-GPU parity check, run under torch.distributed.run with one process per GPU (see
tests/test_gpu_multi.py):  the marker-sharded chain (NCCL all-reduce of the residual deltas every
sync_rate steps) against the CPU oracle with the same total number of virtual ranks and one residual
replica per GPU.  Rank 0 prints MGPU_OK on success."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmrm_b200 import api, synth          # noqa: E402
from oracle import oracle_py as O         # noqa: E402


def chain_check(rank, world, local, sync_rate, Vl=8):
    """The check itself, inside an initialised process group (also called by bench.py before it times N > 1 GPUs).
    Raises on any mismatch; returns a small record on success."""
    N, M, T, G, iters, seed = 3001, 1203, 2, 2, 4, 77
    R = Vl * world
    obj = [None]
    if rank == 0:
        with tempfile.TemporaryDirectory() as tmp:
            d = synth.write_dataset(tmp, N=N, M=M, n_traits=T, n_groups=G, na_rate=0.01, missing_rate=0.004, seed=9)
            p = d["paths"]
            obj = [O.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])]
    dist.broadcast_object_list(obj, src=0)
    inp = obj[0]
    K = inp["cva"].shape[1]
    e = api.Engine(N=N, Mt=M, T=T, G=G, K=K, vranks=R, world_size=world, world_rank=rank, sync_rate=sync_rate, seed=seed, device=local)
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    e.comm_init(uid[0])
    lo, n = e.marker_begin, e.marker_count
    e.upload_bed(inp["bed"][lo:lo + n])
    e.finalize_bed()
    if sync_rate == 1:

        def gather(x):
            out = [None] * world
            dist.all_gather_object(out, x)
            return out
        e.exchange_buffers(gather)
    for t in range(T):
        e.set_phenotype(t, inp["eps0"][t], inp["mask4"][t], int(inp["nonas"][t]))
    e.set_groups(inp["group_index"], inp["cva"])
    e.compute_marker_stats()
    e.init_chain(None)
    hist = []
    for i in range(iters):
        e.run_iteration(i + 1)
        st = e.state()
        st["betas"] = np.stack([e.betas(t) for t in range(T)])
        st["comp"] = np.stack([e.components(t) for t in range(T)])
        hist.append(st)
    eps = np.stack([e.epsilon(t) for t in range(T)])
    e.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, n, hist, eps))
    err = None
    try:
        _compare(inp, gathered, N, R, T, world, sync_rate, iters, seed, rank)
    except AssertionError as ex:
        err = f"rank {rank}: {ex}"[:600]
    errs = [None] * world
    dist.all_gather_object(errs, err)
    errs = [x for x in errs if x]
    if errs:
        raise AssertionError("; ".join(errs))
    return {"sync_rate": sync_rate, "R": R, "world": world, "iterations": iters, "inputs": inp}


def _compare(inp, gathered, N, R, T, world, sync_rate, iters, seed, rank):
    delta_mode = sync_rate > 1 or os.environ.get("GMRM_EXCHANGE") == "delta"   # one residual replica per GPU, deltas all-reduced
    if not delta_mode:          # list exchange: ONE chain, every GPU holds the same residuals bit for bit
        for (_, _, _, eps_g) in gathered[1:]:
            assert np.array_equal(eps_g, gathered[0][3]), "residual replicas are not bit-identical"
    if rank == 0:
        # sync_rate 1: list exchange, every GPU holds the residuals of ONE chain with R virtual ranks (the reference under
        # mpirun -n R); sync_rate > 1: one residual replica per GPU, deltas all-reduced every sync_rate steps
        res = O.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                      nrep=world if delta_mode else 1, iterations=iters, rng_mode=1, seed=seed, sync_rate=sync_rate)
        for (lo_g, n_g, hist_g, eps_g) in gathered:
            for i in range(iters):
                assert np.array_equal(hist_g[i]["comp"], res["comp"][i][:, lo_g:lo_g + n_g]), f"comp differs it {i + 1}"
                np.testing.assert_allclose(hist_g[i]["betas"], res["betas"][i][:, lo_g:lo_g + n_g], rtol=1e-8, atol=1e-13)
                np.testing.assert_allclose(hist_g[i]["sigmag"], res["sigmag"][i], rtol=1e-8)
                np.testing.assert_allclose(hist_g[i]["sigmae"], res["sigmae"][i], rtol=1e-8)
                np.testing.assert_allclose(hist_g[i]["pi"], res["pi"][i], rtol=1e-8)
                assert np.array_equal(hist_g[i]["m0"], res["m0"][i])
        # replicas agree with each other and with the oracle's replica 0 up to rounding
        for (_, _, _, eps_g) in gathered:
            np.testing.assert_allclose(eps_g, res["eps_final"][:, :N], rtol=0, atol=1e-10)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    sync_rate = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    info = chain_check(rank, world, local, sync_rate)
    if len(sys.argv) > 2 and sys.argv[2] == "predict":
        predict_check(rank, world, local, info["inputs"], 3001, 1203, 2)
    if rank == 0:
        print(f"MGPU_OK world={world} sync_rate={sync_rate} R={info['R']}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def predict_check(rank, world, local, inp, N, M, T):
    """Bayes::predict (bayes.cpp:14-284) over the marker shards: one all-reduce of the genetic values, every GPU its
    blocks' statistics; against the oracle with the same total number of blocks.  Rank 0 prints MGPU_PREDICT_OK."""
    Vl = 2
    R = Vl * world
    rng = np.random.default_rng(5)
    hist = rng.normal(0, 0.02, size=(3, M)) * (rng.random((3, M)) < 0.3)
    keep = (rng.random(M) > 0.04).astype(np.uint8)
    e = api.Engine(N=N, Mt=M, T=T, G=1, K=2, vranks=R, world_size=world, world_rank=rank, device=local)
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    e.comm_init(uid[0])
    lo, n = e.marker_begin, e.marker_count
    e.upload_bed(inp["bed"][lo:lo + n])
    e.finalize_bed()
    for t in range(T):
        e.set_phenotype(t, inp["eps0"][t], inp["mask4"][t], int(inp["nonas"][t]))
    e.compute_marker_stats()
    got = [e.predict(t, inp["eps0"][t], hist.mean(axis=0)[lo:lo + n], keep[lo:lo + n]) for t in range(T)]
    e.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, n, got))
    if rank == 0:
        for t in range(T):
            mave, msig = O.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
            want = O.predict(inp["bed"], inp["mask4"][t], int(inp["nonas"][t]), inp["eps0"][t], mave, msig, hist, N=N, R=R, keep=keep)
            scale = np.abs(want["g"]).max()
            for (lo_g, n_g, got_g) in gathered:
                assert np.abs(got_g[t]["g"] - want["g"][:N]).max() <= 1e-11 * scale
                k = keep[lo_g:lo_g + n_g] != 0
                for name in ("beta", "tdist", "se", "pval"):
                    np.testing.assert_allclose(got_g[t][name][k], want[name][lo_g:lo_g + n_g][k], rtol=1e-9, atol=1e-13, err_msg=name)
                    assert np.all(np.isnan(got_g[t][name][~k]))
        print(f"MGPU_PREDICT_OK world={world} R={R}", flush=True)


if __name__ == "__main__":
    sys.exit(main())
