"""CPU: the whole chain from the kernels' SOURCE -- .bed transcode and missing lists, marker statistics, chain start,
and per iteration the intercept draw, residual offset, step table (permutation), group constants, every marker step
(step kernel + sampler kernel + published list), the flush, sum beta^2, sum eps^2 and the global draws -- executed on the
host through tests/emu/cuda_emu.h in the order gmrm_run_iteration launches them (gmrm_b200/csrc/engine.cu), and compared
with the oracle's trajectory (Bayes::process, src/bayes.cpp:318-656) on the same Philox streams over several iterations.

Every __global__ function of the Gibbs path in gmrm_b200/csrc/kernels.cu runs here except the multi-GPU merge and the test
hooks.  Test infrastructure only (a CUDA thread is a std::thread; shapes are tiny); the CUDA build is judged by the GPU tests."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gmrm_b200 import api, synth
from test_predict_kernels_emulated import p

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import asm_to_host  # noqa: E402

SRC = os.path.join(ROOT, "gmrm_b200", "csrc", "kernels.cu")
EMU = os.path.join(ROOT, "tests", "emu")

TAIL = r'''
extern "C" int emu_chain(const uint8_t* plink, int N, int nsm, int Mt, int T, int G, int K, int R, int iters, uint32_t seed, int shuffle,
                         const double* eps0 /*[T][N]*/, const uint8_t* mask4_in /*[T][mbytes]*/, const int32_t* nonas,
                         const int32_t* group, const double* cva, const double* sigmag_init, const int32_t* plan,
                         double* o_betas, int32_t* o_comp, double* o_sigmag, double* o_sigmae, double* o_pi, double* o_mu, int32_t* o_m0,
                         double* o_eps /*[T][N]*/, uint8_t* o_plink /* round trip of the .bed */, double* o_mave, double* o_msig) {
    using namespace gmrm;
    const Layout L = make_layout(N, nsm);
    const int Mm = (Mt + R - 1) / R;
    int32_t err = 0;
    // ---- ingestion (gmrm_upload_bed / gmrm_finalize_bed): transcode, missing counts -> CSR, lists; and the inverse
    std::vector<uint8_t> bed((size_t)Mt * L.col_stride, 0xee);
    std::vector<uint32_t> cnt(Mt, 0), off(Mt + 1, 0);
    emu_launch(EmuDim3((unsigned)((L.col_stride + 255) / 256), Mt), EmuDim3(256), [&] { transcode_kernel(plink, Mt, L, bed.data(), cnt.data()); });
    for (int j = 0; j < Mt; j++) off[j + 1] = off[j] + cnt[j];
    std::vector<uint32_t> midx(std::max<size_t>(off[Mt], 1), 0xffffffffu);
    emu_launch(EmuDim3((Mt + 3) / 4), EmuDim3(128), [&] { fill_missing_kernel(plink, Mt, L, off.data(), midx.data()); });
    emu_launch(EmuDim3((unsigned)((L.mbytes + 255) / 256), Mt), EmuDim3(256), [&] { untranscode_kernel(bed.data(), Mt, L, o_plink); });
    emu_launch(EmuDim3((Mt + 3) / 4), EmuDim3(128), [&] { unmiss_kernel(Mt, L, off.data(), midx.data(), o_plink); });
    // ---- phenotypes (gmrm_set_phenotype): zero-padded residuals and NA bytes
    std::vector<double> eps((size_t)T * L.npad, 0.0);
    std::vector<uint8_t> mask4((size_t)T * L.col_stride, 0);
    for (int t = 0; t < T; t++)
        for (int i = 0; i < N; i++)
            if ((mask4_in[(size_t)t * L.mbytes + i / 4] >> (i % 4)) & 1) {
                eps[(size_t)t * L.npad + i] = eps0[(size_t)t * N + i];
                mask4[(size_t)t * L.col_stride + i / 4] |= (uint8_t)(1u << (i % 4));
            }
    // ---- statistics
    std::vector<uint32_t> na_off, na_idx;
    emu_na_lists(mask4.data(), L.col_stride, T, N, na_off, na_idx);
    emu_launch(EmuDim3(std::min(Mt, 3)), EmuDim3(kStatsThreads), [&] { stats_kernel(bed.data(), Mt, L, mask4.data(), off.data(), midx.data(), nonas, na_off.data(), na_idx.data(), T, o_mave, o_msig, nullptr); });
    // ---- chain start (gmrm_init_chain)
    std::vector<int32_t> mtotgrp(G, 0);
    for (int j = 0; j < Mt; j++) mtotgrp[group[j]]++;
    std::vector<double> sigmag(sigmag_init, sigmag_init + (size_t)T * G), sigmae(T), pi((size_t)T * G * K), cvai((size_t)G * K, 0.0);
    for (int g = 0; g < G; g++) {
        double sum_cva = 0.0;
        for (int j = 1; j < K; j++) { sum_cva += cva[g * K + j]; cvai[g * K + j] = 1.0 / cva[g * K + j]; }
        for (int t = 0; t < T; t++) {
            double* row = &pi[((size_t)t * G + g) * K];
            row[0] = 0.5;
            for (int j = 1; j < K; j++) row[j] = row[0] * cva[g * K + j] / sum_cva;
        }
        for (int t = 0; t < T; t++) if (mtotgrp[g] == 0) sigmag[t * G + g] = 0.0;
    }
    std::vector<double> betas((size_t)T * Mt, 0.0), mu(T, 0.0), mu_old(T, 0.0), esq(T), bsq((size_t)T * G), gc((size_t)T * G * 4 * K);
    std::vector<int32_t> comp((size_t)T * Mt, 0), cass((size_t)T * G * K, 0), m0((size_t)T * G, 0), steptab((size_t)Mm * R);
    emu_launch(EmuDim3(T), EmuDim3(1024), [&] { eps_sumsq_kernel(eps.data(), L.npad, L.npad, esq.data()); });
    emu_launch(EmuDim3(1), EmuDim3(32), [&] { init_sigmae_kernel(esq.data(), nonas, T, sigmae.data()); });

    std::vector<double> partial((size_t)R * T * nsm), spart((size_t)T * nsm), plist((size_t)T * publist_doubles(R), 0.0);
    std::vector<PubEntry> pub((size_t)R * T);
    unsigned long long last_seq = 0;
    int64_t npub = 0;
    const int32_t* cols = nullptr;
    auto step = [&](int V, bool pending, const int32_t* pl) {
        for (int t0 = 0; t0 < T; t0 += pl[0]) {
            StepParams q{};
            q.bed = bed.data(); q.col_stride = L.col_stride; q.nrows = L.nrows; q.cols = cols; q.V = V; q.eps = eps.data(); q.npad = L.npad;
            q.Ttot = T; q.t0 = t0; q.rows_per_pass = pl[1]; q.npass = pl[2]; q.partial = partial.data(); q.spart = spart.data();
            q.mask4 = mask4.data(); q.pV = R; q.err = &err; q.pf = 1;
            if (pending) { q.pG = 1; q.plist = plist.data(); q.wait_seq = last_seq; q.pbed[0] = bed.data(); q.pmiss_off[0] = off.data(); q.pmiss_idx[0] = midx.data(); }
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
    for (int it = 1; it <= iters; it++) {
        // ---- prologue, in the order of gmrm_run_iteration
        MuDrawParams mp{};
        mp.T = T; mp.it = it; mp.seed = seed; mp.sigmae = sigmae.data(); mp.nonas = nonas; mp.mu = mu.data(); mp.mu_old = mu_old.data();
        emu_launch(EmuDim3(1), EmuDim3(32), [&] { mu_draw_kernel(mp); });
        emu_launch(EmuDim3((unsigned)((L.npad + 255) / 256), T), EmuDim3(256), [&] { eps_offset_kernel(eps.data(), mask4.data(), L, mu_old.data(), mu.data()); });
        emu_launch(EmuDim3((unsigned)(((int64_t)Mm * R + 255) / 256)), EmuDim3(256), [&] { steptab_kernel(steptab.data(), Mm, R, 0, R, Mt, 0, shuffle, seed, it, nullptr); });
        emu_launch(EmuDim3((T * G + 127) / 128), EmuDim3(128), [&] {
            group_consts_kernel(T, G, K, N, sigmag.data(), sigmae.data(), pi.data(), cva, cvai.data(), nonas, gc.data());
        });
        std::fill(cass.begin(), cass.end(), 0);
        // ---- marker loop
        for (int s = 0; s < Mm; s++) {
            cols = steptab.data() + (size_t)s * R;
            step(R, s > 0, plan);
            SampleParams sp{};
            sp.V = R; sp.T = T; sp.G = G; sp.K = K; sp.N = N; sp.nsm = nsm; sp.it = it; sp.seed = seed; sp.r0 = 0; sp.R = R; sp.step = s;
            sp.marker_begin = 0; sp.Mloc = Mt; sp.cols = cols; sp.partial = partial.data(); sp.spart = spart.data();
            sp.miss_off = off.data(); sp.miss_idx = midx.data(); sp.eps = eps.data(); sp.npad = L.npad; sp.mave = o_mave; sp.msig = o_msig;
            sp.betas = betas.data(); sp.comp = comp.data(); sp.group = group; sp.sigmag = sigmag.data(); sp.gc = gc.data(); sp.nonas = nonas;
            sp.cass = cass.data(); sp.pub = pub.data(); sp.plist = plist.data(); sp.world = 1; sp.rank = 0;
            sp.seq = (unsigned long long)s + 1; sp.err = &err; sp.npublished = &npub;
            emu_launch(EmuDim3(publist_segments(R)), EmuDim3(kSegCap * 32), [&] { sample_kernel(sp); });
            last_seq = sp.seq;
        }
        cols = nullptr;
        step(0, true, plan + 3);                             // flush
        // ---- epilogue
        emu_launch(EmuDim3(T), EmuDim3(256), [&] { beta_sq_kernel(betas.data(), group, Mt, G, bsq.data()); });
        if (it == 1) {   // the two-level form the engine uses for shards of 4,096 markers and more: same sums
            std::vector<double> part((size_t)T * kBsqSlices * G, -1.0), bsq2((size_t)T * G, -1.0);
            emu_launch(EmuDim3(T, kBsqSlices), EmuDim3(256), [&] { beta_sq_part_kernel(betas.data(), group, Mt, G, part.data()); });
            emu_launch(EmuDim3((T * G + 127) / 128), EmuDim3(128), [&] { beta_sq_final_kernel(part.data(), T, G, bsq2.data()); });
            for (int i = 0; i < T * G; i++)
                if (std::fabs(bsq2[i] - bsq[i]) > 1e-12 * std::max(1.0, std::fabs(bsq[i]))) err = 77;
        }
        emu_launch(EmuDim3(T), EmuDim3(1024), [&] { eps_sumsq_kernel(eps.data(), L.npad, N, esq.data()); });
        GlobalDrawParams gp{};
        gp.T = T; gp.G = G; gp.K = K; gp.N = N; gp.it = it; gp.seed = seed; gp.mtotgrp = mtotgrp.data(); gp.bsq = bsq.data();
        gp.cass = cass.data(); gp.esq = esq.data(); gp.sigmag = sigmag.data(); gp.sigmae = sigmae.data(); gp.pi = pi.data(); gp.m0 = m0.data();
        gp.err = &err;
        emu_launch(EmuDim3(1), EmuDim3(32), [&] { global_draw_kernel(gp); });
        const size_t ih = (size_t)(it - 1);
        memcpy(o_betas + ih * T * Mt, betas.data(), sizeof(double) * T * Mt);
        memcpy(o_comp + ih * T * Mt, comp.data(), sizeof(int32_t) * T * Mt);
        memcpy(o_sigmag + ih * T * G, sigmag.data(), sizeof(double) * T * G);
        memcpy(o_sigmae + ih * T, sigmae.data(), sizeof(double) * T);
        memcpy(o_pi + ih * T * G * K, pi.data(), sizeof(double) * T * G * K);
        memcpy(o_mu + ih * T, mu.data(), sizeof(double) * T);
        memcpy(o_m0 + ih * T * G, m0.data(), sizeof(int32_t) * T * G);
    }
    for (int t = 0; t < T; t++) memcpy(o_eps + (size_t)t * N, eps.data() + (size_t)t * L.npad, sizeof(double) * N);
    return err;
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    text = open(SRC).read()
    names = ("helpers", "ingest", "stats", "eps", "step", "sample", "epilogue")
    parts = [text[text.index(f"// [{n}-begin]"):text.index(f"// [{n}-end]")] for n in names]
    body, _ = asm_to_host.rewrite("".join(parts))
    body = body.replace("#pragma unroll\n", "")
    body = body.replace("extern __shared__ __align__(16) uint8_t smem_raw[];", "uint8_t* smem_raw = emu_smem_storage + 16;")
    assert "extern __shared__ double acc[];" in body
    body = body.replace("extern __shared__ double acc[];", "double* acc = reinterpret_cast<double*>(emu_smem_storage);")
    d = tmp_path_factory.mktemp("emu_chain")
    cpp = d / "chain_emu.cpp"
    cpp.write_text('#include "cuda_emu.h"\n#include "kernels.cuh"\nnamespace gmrm {\n' + body + "\n}\n" + TAIL)
    so = d / "libchain_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                    "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-I", os.path.join(EMU, "fake_cuda"), "-I", EMU,
                    "-I", os.path.join(ROOT, "gmrm_b200", "csrc"), str(cpp), "-o", str(so)], check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("N,M,T,G,R,nsm,iters,shuffle", [(515, 48, 1, 1, 4, 1, 3, 1), (770, 60, 2, 2, 5, 2, 3, 1), (259, 40, 1, 3, 8, 1, 2, 0)])
def test_emulated_chain_matches_oracle(emu, oracle, tmp_path, N, M, T, G, R, nsm, iters, shuffle):
    seed = 977
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=G, na_rate=0.02, missing_rate=0.015, seed=N % 41 + G)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    if G == 3:                                            # an empty group on the way: its sigmaG stays 0, no draws for it
        inp["group_index"] = np.where(inp["group_index"] == 2, 0, inp["group_index"]).astype(np.int32)
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=iters, shuffle=bool(shuffle), rng_mode=1, seed=seed)
    K = inp["cva"].shape[1]
    mbytes = (N + 3) // 4
    p1, p0 = api.step_plan(N, nsm, R, T, want_ranges=False), api.step_plan(N, nsm, 0, T, want_ranges=False)
    plan = np.array([p1["traits_per_launch"], p1["rows_per_pass"], p1["npass"], p0["traits_per_launch"], p0["rows_per_pass"], p0["npass"]], dtype=np.int32)
    bed = np.ascontiguousarray(inp["bed"], dtype=np.uint8)
    eps0 = np.ascontiguousarray(inp["eps0"][:, :N])
    mask4 = np.ascontiguousarray(inp["mask4"][:, :mbytes], dtype=np.uint8)
    nonas = np.ascontiguousarray(inp["nonas"], dtype=np.int32)
    group = np.ascontiguousarray(inp["group_index"], dtype=np.int32)
    cva = np.ascontiguousarray(inp["cva"], dtype=np.float64)
    sg0 = np.ascontiguousarray(res["sigmag_init"], dtype=np.float64)
    o = {"betas": np.zeros((iters, T, M)), "comp": np.zeros((iters, T, M), dtype=np.int32), "sigmag": np.zeros((iters, T, G)),
         "sigmae": np.zeros((iters, T)), "pi": np.zeros((iters, T, G, K)), "mu": np.zeros((iters, T)), "m0": np.zeros((iters, T, G), dtype=np.int32)}
    eps = np.zeros((T, N)); plink = np.zeros_like(bed); mave = np.zeros((T, M)); msig = np.zeros((T, M))
    rc = emu.emu_chain(p(bed), N, nsm, M, T, G, K, R, iters, C.c_uint32(seed), shuffle, p(eps0), p(mask4), p(nonas), p(group), p(cva), p(sg0),
                       p(plan), p(o["betas"]), p(o["comp"]), p(o["sigmag"]), p(o["sigmae"]), p(o["pi"]), p(o["mu"]), p(o["m0"]),
                       p(eps), p(plink), p(mave), p(msig))
    assert rc == 0
    assert np.array_equal(plink, bed)                                       # .bed -> base-3 quads + lists -> .bed: bit-exact
    for t in range(T):
        mave_o, msig_o = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        np.testing.assert_allclose(mave[t], mave_o, rtol=1e-13)
        np.testing.assert_allclose(msig[t], msig_o, rtol=1e-12)
    for i in range(iters):
        assert np.array_equal(o["comp"][i], res["comp"][i]), f"components differ in iteration {i + 1}"
        np.testing.assert_allclose(o["betas"][i], res["betas"][i], rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(o["sigmag"][i], res["sigmag"][i], rtol=1e-8)
        np.testing.assert_allclose(o["sigmae"][i], res["sigmae"][i], rtol=1e-8)
        np.testing.assert_allclose(o["pi"][i].reshape(T, G * K), np.asarray(res["pi"][i]).reshape(T, G * K), rtol=1e-8)
        np.testing.assert_allclose(o["mu"][i], res["mu"][i], rtol=1e-8, atol=1e-15)
        assert np.array_equal(o["m0"][i], res["m0"][i])
    np.testing.assert_allclose(eps, res["eps_final"][:, :N], rtol=0, atol=1e-10)
