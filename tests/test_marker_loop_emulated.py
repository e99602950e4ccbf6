"""CPU: one whole marker loop of the Gibbs path -- step kernel -> sampler kernel -> published list -> next step kernel, for
every step of iteration 1, then the flush -- with the SOURCE of both kernels (gmrm_b200/csrc/kernels.cu) run on the host
through tests/emu/cuda_emu.h, against the oracle's restatement of Bayes::process (src/bayes.cpp:340-560) following the
same Philox streams (and, second variant, replaying the oracle's logged variates through the kernel's replay inputs).

Covers what the per-kernel emulation tests cannot: the hand-over between the kernels (per-CTA partial sums, ordered
compaction of the published updates into list segments whose headers carry the step's sequence number, the pending list applied by the next launch), the
sampler's plumbing (missing-genotype correction of sum b*eps, group constants, component counts).  Test infrastructure
only; shapes are tiny (a CUDA thread is a std::thread here)."""
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
extern "C" int emu_marker_loop(const uint8_t* bed, int N, int nsm, int T, int G, int K, int R, int Mt, int Mm, const int32_t* perm,
                               double* eps, const uint8_t* mask4, const int32_t* nonas, const double* mave, const double* msig,
                               const int32_t* group, const double* sigmag, const double* sigmae, const double* pi, const double* cva,
                               const double* cvai, const double* rep_u, const double* rep_z, uint32_t seed, const uint32_t* miss_off,
                               const uint32_t* miss_idx, const int32_t* plan /* tc, rpp, npass for V = R, then for V = 0 */,
                               double* betas, int32_t* comp, int32_t* cass, int64_t* npublished) {
    using namespace gmrm;
    const Layout L = make_layout(N, nsm);
    int32_t err = 0;
    std::vector<double> gc((size_t)T * G * 4 * K), partial((size_t)R * T * nsm), spart((size_t)T * nsm), plist((size_t)T * publist_doubles(R), 0.0);
    std::vector<PubEntry> pub((size_t)R * T);
    std::vector<int32_t> cols(R);
    unsigned long long last_seq = 0;                         // sequence number in the segment headers of the pending list
    emu_launch(EmuDim3((T * G + 127) / 128), EmuDim3(128), [&] { group_consts_kernel(T, G, K, N, sigmag, sigmae, pi, cva, cvai, nonas, gc.data()); });

    auto step = [&](int V, bool pending, const int32_t* pl) {
        for (int t0 = 0; t0 < T; t0 += pl[0]) {
            StepParams q{};
            q.bed = bed; q.col_stride = L.col_stride; q.nrows = L.nrows; q.cols = cols.data(); q.V = V; q.eps = eps; q.npad = L.npad;
            q.Ttot = T; q.t0 = t0; q.rows_per_pass = pl[1]; q.npass = pl[2]; q.partial = partial.data(); q.spart = spart.data();
            q.mask4 = mask4; q.pV = R; q.err = &err; q.pf = 1;
            if (pending) { q.pG = 1; q.plist = plist.data(); q.wait_seq = last_seq; q.pbed[0] = bed; q.pmiss_off[0] = miss_off; q.pmiss_idx[0] = miss_idx; }
            const int Tl = std::min((int)pl[0], T - t0);
            emu_launch(EmuDim3(nsm), EmuDim3(kStepThreads), [&] {
                switch (Tl) {
                case 1: step_kernel<1>(q); break;
                case 2: step_kernel<2>(q); break;
                case 3: step_kernel<3>(q); break;
                case 4: step_kernel<4>(q); break;
                }
            });
        }
    };
    for (int s = 0; s < Mm; s++) {
        for (int r = 0; r < R; r++) {                       // Bayes::set_block_of_markers (bayes.cpp:903-925) + the rank's permutation
            const int size = Mt / R, modu = Mt % R, Sr = r * size + std::min(r, modu);
            const int loc = perm[(size_t)r * Mm + s];
            cols[r] = loc >= 0 ? Sr + loc : -1;
        }
        step(R, s > 0, plan);
        SampleParams sp{};
        sp.V = R; sp.T = T; sp.G = G; sp.K = K; sp.N = N; sp.nsm = nsm; sp.it = 1; sp.seed = seed; sp.r0 = 0; sp.R = R; sp.step = s;
        sp.marker_begin = 0; sp.Mloc = Mt; sp.cols = cols.data(); sp.partial = partial.data(); sp.spart = spart.data();
        sp.miss_off = miss_off; sp.miss_idx = miss_idx; sp.eps = eps; sp.npad = L.npad; sp.mave = mave; sp.msig = msig;
        sp.betas = betas; sp.comp = comp; sp.group = group; sp.sigmag = sigmag; sp.gc = gc.data(); sp.nonas = nonas; sp.cass = cass;
        sp.pub = pub.data(); sp.plist = plist.data(); sp.world = 1; sp.rank = 0; sp.seq = (unsigned long long)s + 1;
        sp.rep_u = rep_u; sp.rep_z = rep_z; sp.err = &err; sp.npublished = npublished;
        emu_launch(EmuDim3(publist_segments(R)), EmuDim3(kSegCap * 32), [&] { sample_kernel(sp); });
        last_seq = sp.seq;
    }
    step(0, true, plan + 3);                                 // flush: the last step's updates
    return err;
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    text = open(SRC).read()
    parts = [text[text.index(f"// [{n}-begin]"):text.index(f"// [{n}-end]")] for n in ("helpers", "step", "sample")]
    body, n = asm_to_host.rewrite("".join(parts))
    body = body.replace("#pragma unroll\n", "")
    decl = "extern __shared__ __align__(16) uint8_t smem_raw[];"
    body = body.replace(decl, "uint8_t* smem_raw = emu_smem_storage + 16;")
    d = tmp_path_factory.mktemp("emu_loop")
    cpp = d / "loop_emu.cpp"
    cpp.write_text('#include "cuda_emu.h"\n#include "kernels.cuh"\nnamespace gmrm {\n' + body + "\n}\n" + TAIL)
    so = d / "libloop_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                    "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-I", os.path.join(EMU, "fake_cuda"), "-I", EMU,
                    "-I", os.path.join(ROOT, "gmrm_b200", "csrc"), str(cpp), "-o", str(so)], check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("N,M,T,G,R,nsm,replay", [(515, 60, 1, 1, 4, 1, False), (1030, 90, 2, 2, 6, 2, False), (515, 60, 2, 2, 5, 1, True)])
def test_emulated_marker_loop_matches_oracle(emu, oracle, tmp_path, N, M, T, G, R, nsm, replay):
    seed = 4242
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=G, na_rate=0.02, missing_rate=0.01, seed=N % 53 + R)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=1, rng_mode=1, seed=seed)
    K = inp["cva"].shape[1]
    Mm = -(-M // R)
    tri, miss_off, miss_idx, nrows = to_device_layout(inp["bed"], N, nsm)
    npad, stride = nrows * 256, nrows * 64
    mask4 = np.zeros((T, stride), dtype=np.uint8)
    mask4[:, : inp["mask4"].shape[1]] = inp["mask4"]
    obs = np.stack([((inp["mask4"][t][:, None] >> np.arange(4)) & 1).reshape(-1)[:N].astype(bool) for t in range(T)])
    # state at the start of iteration 1's marker loop (bayes.cpp:322-335, 347-368; phenotype.cpp:432-459)
    eps = np.zeros((T, npad))
    eps[:, :N] = inp["eps0"][:, :N]
    sigmae = np.array([(eps[t, :N] ** 2 * obs[t]).sum() / int(inp["nonas"][t]) * 0.5 for t in range(T)])
    for t in range(T):
        eps[t, :N] -= res["mu_draw"][0][t] * obs[t]
    sigmag = np.ascontiguousarray(res["sigmag_init"], dtype=np.float64)
    cva = np.ascontiguousarray(inp["cva"], dtype=np.float64)
    cvai = np.zeros_like(cva)
    cvai[:, 1:] = 1.0 / cva[:, 1:]                                                         # options.cpp:282
    pi = np.full((T, G, K), 0.5)
    for g in range(G):
        pi[:, g, 1:] = 0.5 * cva[g, 1:] / cva[g, 1:].sum()                                 # bayes.hpp:34-47
    mave = np.empty((T, M)); msig = np.empty((T, M))
    for t in range(T):
        mave[t], msig[t] = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
    perm = np.ascontiguousarray(res["perm"][0], dtype=np.int32)                            # [R][Mm]
    rep_u = np.ascontiguousarray(res["u"][0]) if replay else None
    rep_z = np.ascontiguousarray(res["z"][0]) if replay else None
    p1, p0 = api.step_plan(N, nsm, R, T, want_ranges=False), api.step_plan(N, nsm, 0, T, want_ranges=False)
    plan = np.array([p1["traits_per_launch"], p1["rows_per_pass"], p1["npass"], p0["traits_per_launch"], p0["rows_per_pass"], p0["npass"]], dtype=np.int32)
    betas = np.zeros((T, M)); comp = np.zeros((T, M), dtype=np.int32); cass = np.zeros((T, G * K), dtype=np.int32)
    npub = np.zeros(1, dtype=np.int64)
    nonas = np.ascontiguousarray(inp["nonas"], dtype=np.int32)
    group = np.ascontiguousarray(inp["group_index"], dtype=np.int32)
    rc = emu.emu_marker_loop(p(tri), N, nsm, T, G, K, R, M, Mm, p(perm), p(eps), p(mask4), p(nonas), p(mave), p(msig), p(group),
                             p(sigmag), p(sigmae), p(pi), p(cva), p(cvai), p(rep_u), p(rep_z), C.c_uint32(seed), p(miss_off), p(miss_idx),
                             p(plan), p(betas), p(comp), p(cass), p(npub))
    assert rc == 0
    assert np.array_equal(comp, res["comp"][0])                                            # integer work: bit-exact
    np.testing.assert_allclose(betas, res["betas"][0], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(eps[:, :N], res["eps_final"][:, :N], rtol=0, atol=1e-11)
    assert not eps[:, N:].any()
    mtot = np.bincount(group, minlength=G)
    for t in range(T):
        assert np.array_equal(mtot - cass[t].reshape(G, K)[:, 0], res["m0"][0][t])         # bayes.cpp:605
        assert cass[t].sum() == M
    assert npub[0] > 0                                                                     # updates were published and applied
