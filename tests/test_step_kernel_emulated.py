"""CPU: the step kernel of the Gibbs path (gmrm_b200/csrc/kernels.cu: pending updates -> look-up tables -> column
stream) executed on the host through tests/emu/cuda_emu.h, its inline PTX rewritten statement by statement
(tests/emu/asm_to_host.py: PRMT, LDS/STS with absolute shared addresses, streaming loads), and compared with the oracle:
Bayes::dot_product's sums (src/bayes.cpp:749-766) and Phenotype::update_epsilon (src/phenotype.cpp:326-390).

One std::thread per CUDA thread (512 per CTA), barriers for __syncthreads and the warp shuffles.  It checks the
kernel's logic -- table geometry, row ownership, the transposed butterfly, the published-list update with missing
genotypes and NAs -- on shapes that take seconds; it says nothing about the CUDA build or its speed, which stay with
the GPU parity tests.  Test infrastructure only: the product has no CPU path."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gmrm_b200 import api, synth
from test_predict_kernels_emulated import p, to_device_layout

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import asm_to_host  # noqa: E402

SRC = os.path.join(ROOT, "gmrm_b200", "csrc", "kernels.cu")
EMU = os.path.join(ROOT, "tests", "emu")

TAIL = r'''
extern "C" int emu_step(const uint8_t* bed, int N, int nsm, const int32_t* cols, int V, double* eps, const uint8_t* mask4, int T,
                        int tc, int rpp, int npass, double* partial, double* spart, const double* plist, int pV,
                        const uint32_t* miss_off, const uint32_t* miss_idx, const uint8_t* bed2) {
    using namespace gmrm;
    const Layout L = make_layout(N, nsm);
    int32_t err = 0;
    for (int t0 = 0; t0 < T; t0 += tc) {
        StepParams q{};
        if (bed2 && T == 1) { q.bed2 = bed2; q.drows = npass * nsm; q.ndir = 1; }   // hybrid plan: last row of a pass is a direct row
        q.bed = bed; q.col_stride = L.col_stride; q.nrows = L.nrows; q.cols = cols; q.V = V; q.eps = eps; q.npad = L.npad;
        q.Ttot = T; q.t0 = t0; q.rows_per_pass = rpp; q.npass = npass; q.partial = partial; q.spart = spart;
        q.mask4 = mask4; q.pV = pV; q.err = &err; q.pf = 1;
        if (plist) { q.pG = 1; q.plist = plist; q.pbed[0] = bed; q.pmiss_off[0] = miss_off; q.pmiss_idx[0] = miss_idx; }
        const int Tl = std::min(tc, T - t0);
        emu_launch(EmuDim3(nsm), EmuDim3(kStepThreads), [&] {
            switch (Tl) {
            case 1: step_kernel<1>(q); break;
            case 2: step_kernel<2>(q); break;
            case 3: step_kernel<3>(q); break;
            case 4: step_kernel<4>(q); break;
            }
        });
    }
    return err;
}

// update-only launch with the lists of TWO GPUs (world_size 2, list exchange): list g holds columns of shard g, whose
// bytes and missing lists are read from "GPU g's" buffers (peer memory on hardware); this GPU is shard 0.
// rs_world > 0: the row-sharded form -- the launch is repeated for every "GPU" rs_rank = 0 .. rs_world-1 in turn on the SAME
// residual array (each updates only its rows of every CTA and "stores them to every GPU"); the row flags are preset, nothing waits
extern "C" int emu_update_two_lists(const uint8_t* bed0, const uint8_t* bed1, int N, int nsm, double* eps, const uint8_t* mask4, int T,
                                    int tc, int rpp, int npass, const double* plists /* [2][T][ld] */, int pV,
                                    const uint32_t* moff0, const uint32_t* midx0, const uint32_t* moff1, const uint32_t* midx1, int rs_world) {
    using namespace gmrm;
    const Layout L = make_layout(N, nsm);
    int32_t err = 0;
    std::vector<unsigned long long> flags_set((size_t)kMaxGpus * nsm, 1000ull), flags_out((size_t)kMaxGpus * nsm, 0ull);
    for (int rs_rank = 0; rs_rank < std::max(rs_world, 1); rs_rank++)
    for (int t0 = 0; t0 < T; t0 += tc) {
        StepParams q{};
        if (rs_world > 0) {
            q.rs_world = rs_world; q.rs_rank = rs_rank; q.row_seq = 7; q.rflag_mine = flags_set.data();
            for (int g = 0; g < rs_world; g++) { q.peps[g] = eps; q.rflag_peer[g] = flags_out.data(); }
        }
        q.bed = bed0; q.col_stride = L.col_stride; q.nrows = L.nrows; q.V = 0; q.eps = eps; q.npad = L.npad;
        q.Ttot = T; q.t0 = t0; q.rows_per_pass = rpp; q.npass = npass; q.mask4 = mask4; q.pV = pV; q.err = &err; q.pf = 1;
        q.pG = 2; q.plist = plists;
        q.pbed[0] = bed0; q.pmiss_off[0] = moff0; q.pmiss_idx[0] = midx0;
        q.pbed[1] = bed1; q.pmiss_off[1] = moff1; q.pmiss_idx[1] = midx1;
        const int Tl = std::min(tc, T - t0);
        emu_launch(EmuDim3(nsm), EmuDim3(kStepThreads), [&] {
            switch (Tl) {
            case 1: step_kernel<1>(q); break;
            case 2: step_kernel<2>(q); break;
            case 3: step_kernel<3>(q); break;
            case 4: step_kernel<4>(q); break;
            }
        });
    }
    return err;
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    text = open(SRC).read()
    helpers = text[text.index("// [helpers-begin]"):text.index("// [helpers-end]")]
    step = text[text.index("// [step-begin]"):text.index("// [step-end]")]
    body, n = asm_to_host.rewrite(helpers + step)
    assert n >= 14                                              # every inline-PTX statement of the region has a host form
    body = body.replace("#pragma unroll\n", "")
    decl = "extern __shared__ __align__(16) uint8_t smem_raw[];"
    assert decl in body
    body = body.replace(decl, "uint8_t* smem_raw = emu_smem_storage + 16;")
    d = tmp_path_factory.mktemp("emu_step")
    cpp = d / "step_emu.cpp"
    cpp.write_text('#include "cuda_emu.h"\n#include "kernels.cuh"\nnamespace gmrm {\n' + body + "\n}\n" + TAIL)
    so = d / "libstep_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                    "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-I", os.path.join(EMU, "fake_cuda"), "-I", EMU,
                    "-I", os.path.join(ROOT, "gmrm_b200", "csrc"),
                    # the build variant with the hybrid plan compiled in (12 warps, one direct row per pass); the cases without a
                    # second genotype copy run the look-up-only path of the same source, as the product build (16 warps) does
                    "-DGMRM_STEP_WARPS=12", "-DGMRM_STEP_DIRECT=1", str(cpp), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def dosages(bed, N):
    codes = (bed[:, :, None] >> (2 * np.arange(4))) & 3
    return np.where(codes == 0, 2.0, np.where(codes == 2, 1.0, 0.0)).reshape(bed.shape[0], -1)[:, :N]


@pytest.mark.parametrize("N,nsm,T,M,na,miss,hybrid", [(777, 1, 3, 33, 0.02, 0.01, False), (1795, 2, 2, 40, 0.0, 0.02, False), (5119, 1, 1, 50, 0.01, 0.0, False),
                                                      (3000, 3, 1, 21, 0.03, 0.03, False), (1024, 1, 4, 18, 0.0, 0.0, False),
                                                      (5119, 1, 1, 50, 0.01, 0.0, True), (3000, 3, 1, 21, 0.03, 0.03, True), (8000, 2, 1, 19, 0.02, 0.01, True)])
def test_emulated_step_kernel_matches_oracle(emu, oracle, tmp_path, N, nsm, T, M, na, miss, hybrid):
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=1, na_rate=na, missing_rate=miss, seed=N % 71)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    tri, miss_off, miss_idx, nrows = to_device_layout(inp["bed"], N, nsm)
    npad, stride = nrows * 256, nrows * 64
    mask4 = np.zeros((T, stride), dtype=np.uint8)
    mask4[:, : inp["mask4"].shape[1]] = inp["mask4"]
    eps = np.zeros((T, npad))
    eps[:, :N] = inp["eps0"][:, :N]
    a = dosages(inp["bed"], N)
    rng = np.random.default_rng(M)

    def launch(cols, plist=None, pV=0):
        V = len(cols)
        plan = api.step_plan(N, nsm, V, T, want_ranges=hybrid)
        assert plan is not None
        bed2 = None
        if hybrid and V:             # second copy of the direct rows (engine: direct_plane_kernel): 2-bit dosage fields, missing = 0
            codes = (inp["bed"][:, :, None] >> (2 * np.arange(4))) & 3
            fields = (np.where(codes == 0, 2, np.where(codes == 2, 1, 0)) << (2 * np.arange(4))).sum(axis=2).astype(np.uint8)
            padded = np.zeros((M, stride), dtype=np.uint8)
            padded[:, : fields.shape[1]] = fields
            npass = plan["npass"]
            bed2 = np.zeros((M, npass * nsm, 64), dtype=np.uint8)
            for q in range(npass):
                for c in range(nsm):
                    start, count = plan["ranges"][q][c]
                    if count >= 2:
                        bed2[:, q * nsm + c, :] = padded[:, (start + count - 1) * 64: (start + count) * 64]
        cols_a = np.ascontiguousarray(cols, dtype=np.int32) if V else np.zeros(1, np.int32)
        partial = np.full((max(V, 1), T, nsm), np.nan)
        spart = np.full((T, nsm), np.nan)
        rc = emu.emu_step(p(tri), N, nsm, p(cols_a), V, p(eps), p(mask4), T, plan["traits_per_launch"], plan["rows_per_pass"],
                          plan["npass"], p(partial), p(spart), p(plist), pV, p(miss_off), p(miss_idx), p(bed2))
        assert rc == 0
        return partial, spart

    # (1) dot products of a shuffled subset of the columns, with a hole (-1 = no marker for that virtual rank)
    cols = rng.permutation(M)[: max(3, (2 * M) // 3)].astype(np.int32)
    cols[1] = -1
    partial, spart = launch(cols)
    for t in range(T):
        got = partial[:, t, :].sum(axis=1)
        want = a[np.maximum(cols, 0)] @ eps[t, :N]
        ok = cols >= 0
        assert np.abs(got[ok] - want[ok]).max() <= 1e-12 * max(np.abs(want).max(), 1.0)
        assert abs(spart[t].sum() - eps[t].sum()) <= 1e-12 * max(np.abs(eps[t]).sum(), 1.0)

    # (2) a published list of the previous step (ordered, compacted: header count + items {lam, mave, col, v}), then dots again
    pV = 37                                                             # three segments, the last one short
    ld = ((pV + 15) // 16) * 50
    plist = np.zeros(T * ld)
    npub = [0] * T
    want_eps = [np.concatenate([eps[t, :N], np.zeros((-N) % 4)]) for t in range(T)]
    for t in range(T):
        mave, msig = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        k = int(rng.integers(1, min(pV, M) + 1)) if t != 1 else 0       # trait 1 publishes nothing
        pub = rng.choice(M, size=k, replace=False)
        vrs = np.sort(rng.choice(pV, size=k, replace=False))            # the virtual ranks that published, in order
        entries = []
        for v, j in zip(vrs, pub):
            db = float(rng.normal(0, 0.05))
            entries.append((int(v), db * msig[j], mave[j], int(j)))
            oracle.update_eps(want_eps[t], inp["mask4"][t], inp["bed"][j], db, mave[j], msig[j])
        plist[t * ld: (t + 1) * ld] = build_list(pV, entries)
        npub[t] = k
    cols2 = rng.permutation(M)[:9].astype(np.int32)
    partial, spart = launch(cols2, plist, pV)
    for t in range(T):
        np.testing.assert_allclose(eps[t, :N], want_eps[t][:N], rtol=0, atol=1e-13)
        assert not eps[t, N:].any()
        got = partial[:, t, :].sum(axis=1)
        want = a[cols2] @ want_eps[t][:N]
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1.0)

    # (3) update-only launch (V = 0): the flush at the end of an iteration
    before = eps.copy()
    launch([], plist, pV)
    for t in range(T):
        changed = np.abs(eps[t] - before[t]).max()
        assert (changed > 0) == (npub[t] > 0)


def build_list(pV, entries):
    """One GPU's published list of one trait in the kernels' segmented layout (kernels.cuh): ceil(pV/16) segments of
    50 doubles (int32 count, pad, 16 items {lam, mave, int32 col, int32 v}); entries = [(virtual rank, lam, mave, col)]."""
    nseg = (pV + 15) // 16
    out = np.zeros(nseg * 50)
    for (v, lam, mave, col) in sorted(entries):
        seg = out[(v // 16) * 50: (v // 16 + 1) * 50]
        i = int(seg[:1].view(np.int32)[0])
        seg[2 + 3 * i], seg[3 + 3 * i] = lam, mave
        seg[4 + 3 * i: 5 + 3 * i].view(np.int32)[:] = (int(col), int(v))
        seg[:1].view(np.int32)[0] = i + 1
    return out


@pytest.mark.parametrize("nsm", [2, 7, 40])
def test_emulated_update_with_the_lists_of_two_gpus(emu, oracle, tmp_path, nsm):
    """world_size 2, list exchange (bayes.cpp:495-553 replaced by published lists): the update phase applies GPU 0's list, then
    GPU 1's -- global virtual-rank order -- reading every published column from the shard that owns it."""
    N, M, T, pV = 2890, 48, 2, 21                              # 12 rows: 6 per CTA, or 1-2 (nsm = 7)
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=1, na_rate=0.02, missing_rate=0.02, seed=12)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    half = M // 2
    shards = [to_device_layout(inp["bed"][:half], N, nsm), to_device_layout(inp["bed"][half:], N, nsm)]
    nrows = shards[0][3]
    npad, stride = nrows * 256, nrows * 64
    mask4 = np.zeros((T, stride), dtype=np.uint8)
    mask4[:, : inp["mask4"].shape[1]] = inp["mask4"]
    eps = np.zeros((T, npad))
    eps[:, :N] = inp["eps0"][:, :N]
    rng = np.random.default_rng(3)
    ld = ((pV + 15) // 16) * 50
    plists = np.zeros((2, T, ld))
    want = [np.concatenate([eps[t, :N], np.zeros((-N) % 4)]) for t in range(T)]
    for t in range(T):
        mave, msig = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        for g in range(2):                                         # GPU order = virtual-rank order
            k = int(rng.integers(1, pV + 1))
            pub = rng.choice(half, size=k, replace=False)           # columns local to shard g
            vrs = np.sort(rng.choice(pV, size=k, replace=False))
            entries = []
            for v, jl in zip(vrs, pub):
                j = g * half + int(jl)
                db = float(rng.normal(0, 0.05))
                entries.append((int(v), db * msig[j], mave[j], int(jl)))
                oracle.update_eps(want[t], inp["mask4"][t], inp["bed"][j], db, mave[j], msig[j])
            plists[g, t] = build_list(pV, entries)
    plan = api.step_plan(N, nsm, 0, T, want_ranges=False)
    eps0 = eps.copy()
    for rs_world in (0, 2, 3, 8):        # every GPU applies everything / the rows of each CTA are shared by 2, 3, 8 GPUs (8: some get none)
        eps[...] = eps0
        rc = emu.emu_update_two_lists(p(shards[0][0]), p(shards[1][0]), N, nsm, p(eps), p(mask4), T, plan["traits_per_launch"],
                                      plan["rows_per_pass"], plan["npass"], p(plists), pV, p(shards[0][1]), p(shards[0][2]),
                                      p(shards[1][1]), p(shards[1][2]), rs_world)
        assert rc == 0
        for t in range(T):
            np.testing.assert_allclose(eps[t, :N], want[t][:N], rtol=0, atol=1e-13)
            assert not eps[t, N:].any()


@pytest.mark.parametrize("N,nsm,T,V", [(1795, 2, 1, 2100), (777, 3, 2, 1100)])
def test_emulated_step_kernel_with_partial_sums_in_global_memory(emu, oracle, tmp_path, N, nsm, T, V):
    """More markers per step than shared memory holds partial sums for (V * T > 2,048): the CTAs accumulate in their slots of
    the global partial array, pass after pass (first pass stores, later passes add); the third CTA of the second case owns no
    rows in some passes.  Dots against the oracle's decode, holes included."""
    M = 40
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=1, na_rate=0.01, missing_rate=0.01, seed=7)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    tri, miss_off, miss_idx, nrows = to_device_layout(inp["bed"], N, nsm)
    npad, stride = nrows * 256, nrows * 64
    mask4 = np.zeros((T, stride), dtype=np.uint8)
    mask4[:, : inp["mask4"].shape[1]] = inp["mask4"]
    eps = np.zeros((T, npad))
    eps[:, :N] = inp["eps0"][:, :N]
    a = dosages(inp["bed"], N)
    rng = np.random.default_rng(V)
    cols = rng.integers(0, M, size=V).astype(np.int32)
    cols[rng.integers(0, V, size=5)] = -1
    plan = api.step_plan(N, nsm, V, T, want_ranges=False)
    assert plan is not None and plan["traits_per_launch"] == T
    partial = np.full((V, T, nsm), np.nan)
    spart = np.full((T, nsm), np.nan)
    rc = emu.emu_step(p(tri), N, nsm, p(cols), V, p(eps), p(mask4), T, plan["traits_per_launch"], plan["rows_per_pass"],
                      plan["npass"], p(partial), p(spart), None, 0, p(miss_off), p(miss_idx), None)
    assert rc == 0
    ok = cols >= 0
    for t in range(T):
        got = partial[:, t, :].sum(axis=1)
        want = a[np.maximum(cols, 0)] @ eps[t, :N]
        assert np.isfinite(got[ok]).all()
        assert np.abs(got[ok] - want[ok]).max() <= 1e-12 * max(np.abs(want).max(), 1.0)
