"""CPU, world_size 2 over gloo: the marker-sharded chain with list exchange (what replaces the reference's
Allgather(bool) + Allgatherv(dbetas) + Allgatherv(column) of bayes.cpp:495-553, and its Allreduce of sum beta^2 / cass,
575-588) with each rank running the SOURCE of the product's kernels through tests/emu/cuda_emu.h for its shard of the
markers, the published lists all-gathered between the launches by torch.distributed (gloo), and every rank applying every
rank's list itself -- reading the other shard's columns where the hardware path reads peer memory.  Both ranks' betas and
components, and the replicated residuals and global parameters, must equal the oracle run with the same total number of
virtual ranks.  Test infrastructure only; the NVLink / NCCL transport itself is covered by tests/test_gpu_multi.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CASE = dict(N=515, M=57, T=2, G=2, Vl=3, nsm=2, iters=2, seed=31)


def build(outdir):
    sys.path.insert(0, EMU)
    import asm_to_host
    text = open(os.path.join(ROOT, "gmrm_b200", "csrc", "kernels.cu")).read()
    names = ("helpers", "ingest", "stats", "eps", "step", "sample", "epilogue")
    body, _ = asm_to_host.rewrite("".join(text[text.index(f"// [{n}-begin]"):text.index(f"// [{n}-end]")] for n in names))
    body = body.replace("#pragma unroll\n", "")
    body = body.replace("extern __shared__ __align__(16) uint8_t smem_raw[];", "uint8_t* smem_raw = emu_smem_storage + 16;")
    body = body.replace("extern __shared__ double acc[];", "double* acc = reinterpret_cast<double*>(emu_smem_storage);")
    cpp = os.path.join(outdir, "chain_mg_emu.cpp")
    with open(cpp, "w") as f:
        f.write('#include "cuda_emu.h"\n#include "kernels.cuh"\nnamespace gmrm {\n' + body + "\n}\n" + open(os.path.join(EMU, "chain_mg_tail.inc")).read())
    so = os.path.join(outdir, "libchain_mg_emu.so")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                    "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-I", os.path.join(EMU, "fake_cuda"), "-I", EMU,
                    "-I", os.path.join(ROOT, "gmrm_b200", "csrc"), cpp, "-o", so], check=True)
    return so


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def worker(rank, world, port, so, data_dir, out, xd=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmrm_b200 import api
    from oracle import oracle_py as O
    c = CASE
    N, M, T, G, Vl, nsm, iters, seed = (c[k] for k in ("N", "M", "T", "G", "Vl", "nsm", "iters", "seed"))
    R = Vl * world
    paths = {k: os.path.join(data_dir, f"syn.{k}") for k in ("bed", "dim", "gri", "grm")}
    inp = O.load_inputs(paths["bed"], paths["dim"], [os.path.join(data_dir, f"syn_t{t}.phen") for t in range(T)], paths["gri"], paths["grm"])
    K = inp["cva"].shape[1]
    res = O.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R, nrep=1,
                  iterations=iters, rng_mode=1, seed=seed)
    lib = C.CDLL(so)
    lib.emu_mg_create.restype = C.c_void_p
    p1, p0 = api.step_plan(N, nsm, Vl, T, want_ranges=False), api.step_plan(N, nsm, 0, T, want_ranges=False)
    plan = np.array([p1["traits_per_launch"], p1["rows_per_pass"], p1["npass"], p0["traits_per_launch"], p0["rows_per_pass"], p0["npass"]], dtype=np.int32)
    mbytes = (N + 3) // 4
    bed = np.ascontiguousarray(inp["bed"], dtype=np.uint8)
    eps0 = np.ascontiguousarray(inp["eps0"][:, :N]); mask4 = np.ascontiguousarray(inp["mask4"][:, :mbytes], dtype=np.uint8)
    nonas = np.ascontiguousarray(inp["nonas"], dtype=np.int32); group = np.ascontiguousarray(inp["group_index"], dtype=np.int32)
    cva = np.ascontiguousarray(inp["cva"], dtype=np.float64); sg0 = np.ascontiguousarray(res["sigmag_init"], dtype=np.float64)
    h = C.c_void_p(lib.emu_mg_create(p(bed), world, rank, N, nsm, M, T, G, K, R, C.c_uint32(seed), p(eps0), p(mask4), p(nonas), p(group),
                                     p(cva), p(sg0), p(plan)))
    mb, ml, mm, ld = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.emu_mg_info(h, C.byref(mb), C.byref(ml), C.byref(mm), C.byref(ld))
    lo, n, Mm, ld = mb.value, ml.value, mm.value, ld.value
    ok = True
    if xd:       # fused increment exchange: residuals, receive buffers and flags of both "GPUs" in one shared mapping
        lib.emu_mg_xd_block_bytes.restype = C.c_size_t
        blk = lib.emu_mg_xd_block_bytes(h)
        shm = np.memmap(os.path.join(data_dir, "xd_shared.bin"), dtype=np.uint8, mode="r+", shape=(world * blk,))
        lib.emu_mg_enable_xd(h, p(shm))
        dist.barrier()
    for it in range(1, iters + 1):
        lib.emu_mg_begin_iteration(h, it)
        lists = None
        for s in range(Mm):
            own = np.zeros(ld)
            ok &= lib.emu_mg_step(h, it, s, p(lists), p(own)) == 0
            if xd:                                                          # the kernels exchange increments themselves: own list only
                lists = own
                continue
            got = [torch.zeros(ld, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(got, torch.from_numpy(own))                     # the exchange: every GPU's list to every GPU
            lists = np.ascontiguousarray(np.stack([g.numpy() for g in got]))
        ok &= lib.emu_mg_flush(h, p(lists)) == 0
        bsq = np.zeros(T * G); cass = np.zeros(T * G * K, dtype=np.int32)
        lib.emu_mg_local_sums(h, p(bsq), p(cass))
        tb, tc = torch.from_numpy(bsq), torch.from_numpy(cass)
        dist.all_reduce(tb); dist.all_reduce(tc)                            # bayes.cpp:575-588
        ok &= lib.emu_mg_global_draw(h, it, p(tb.numpy()), p(tc.numpy())) == 0
        betas = np.zeros((T, n)); comp = np.zeros((T, n), dtype=np.int32); sigmag = np.zeros((T, G)); sigmae = np.zeros(T)
        pi = np.zeros((T, G * K)); mu = np.zeros(T); m0 = np.zeros((T, G), dtype=np.int32); eps = np.zeros((T, N))
        lib.emu_mg_state(h, p(betas), p(comp), p(sigmag), p(sigmae), p(pi), p(mu), p(m0), p(eps))
        i = it - 1
        ok &= bool(np.array_equal(comp, res["comp"][i][:, lo:lo + n]))
        ok &= bool(np.allclose(betas, res["betas"][i][:, lo:lo + n], rtol=1e-8, atol=1e-13))
        ok &= bool(np.allclose(sigmag, res["sigmag"][i], rtol=1e-8)) and bool(np.allclose(sigmae, res["sigmae"][i], rtol=1e-8))
        ok &= bool(np.allclose(pi, np.asarray(res["pi"][i]).reshape(T, G * K), rtol=1e-8)) and bool(np.array_equal(m0, res["m0"][i]))
    ok &= bool(np.allclose(eps, res["eps_final"][:, :N], rtol=0, atol=1e-10))  # the residuals are replicated: every rank holds the chain's
    if xd:                                                                   # one chain: the replicas are bit-identical
        reps = [None] * world
        dist.all_gather_object(reps, eps.tobytes())
        ok &= all(r == reps[0] for r in reps)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, n))                                   # the shards tile the markers in rank order
    ok &= spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1)) and sum(x[1] for x in spans) == M
    out[rank] = bool(ok) and n > 0
    lib.emu_mg_destroy(h)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_chain_with_increment_exchange_matches_oracle(tmp_path):
    """The default multi-GPU exchange (StepParams::xd_world): each rank applies its own list as increments, the step kernel's
    source reduces them sub-slice by sub-slice and stores the new residuals into both ranks' arrays (shared memory stands for
    NVLink peer memory; CTA c of both processes runs concurrently, flags and all)."""
    test_two_rank_chain_with_list_exchange_matches_oracle(tmp_path, xd=True)


def test_three_rank_chain_with_increment_exchange_matches_oracle(tmp_path):
    """The same with three ranks: sub-slices that do not divide the CTA's quads evenly, three increments added in rank order."""
    test_two_rank_chain_with_list_exchange_matches_oracle(tmp_path, xd=True, world=3)


def test_two_rank_chain_with_list_exchange_matches_oracle(tmp_path, xd=False, world=2):
    sys.path.insert(0, ROOT)
    from gmrm_b200 import synth
    c = CASE
    synth.write_dataset(str(tmp_path), N=c["N"], M=c["M"], n_traits=c["T"], n_groups=c["G"], na_rate=0.02, missing_rate=0.015, seed=19)
    so = build(str(tmp_path))
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29000 + os.getpid() % 2000
    if xd:
        with open(os.path.join(str(tmp_path), "xd_shared.bin"), "wb") as f:
            f.write(b"\0" * (64 << 20))
    procs = [ctx.Process(target=worker, args=(r, world, port, so, str(tmp_path), out, xd)) for r in range(world)]
    for q in procs:
        q.start()
    for q in procs:
        q.join(600)
    assert all(q.exitcode == 0 for q in procs), [q.exitcode for q in procs]
    assert out.get(0) is True and out.get(1) is True, dict(out)
