"""CPU: gmrm_predict's launch sequence (gmrm_b200/csrc/engine.cu: statistics with the sum of squares, total genetic
values block by block, then per block g_k, y_k = y - (g - g_k), its variance, the step kernel with y_k as the residual,
the finishing kernel) replayed from the kernels' SOURCE through tests/emu/cuda_emu.h and compared with the oracle's
restatement of Bayes::predict (src/bayes.cpp:14-284) -- including the shapes of tests/test_gpu_predict.py that have not
run on hardware yet (several CTAs, blocks of a single marker).  Test infrastructure only."""
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

CSRC = os.path.join(ROOT, "gmrm_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")

TAIL = r'''
extern "C" int emu_predict(const uint8_t* plink, int N, int nsm, int Mt, int R, const double* y, const uint8_t* mask4_in, int nonas,
                           const double* beta_mean, const uint8_t* keep, int vchunk, int rpp, int npass,
                           double* g_out, double* beta, double* tdist, double* se, double* pval) {
    using namespace gmrm;
    const Layout L = make_layout(N, nsm);
    int32_t err = 0;
    std::vector<uint8_t> bed((size_t)Mt * L.col_stride, 0);
    std::vector<uint32_t> cnt(Mt, 0), off(Mt + 1, 0);
    emu_launch(EmuDim3((unsigned)((L.col_stride + 255) / 256), Mt), EmuDim3(256), [&] { transcode_kernel(plink, Mt, L, bed.data(), cnt.data()); });
    for (int j = 0; j < Mt; j++) off[j + 1] = off[j] + cnt[j];
    std::vector<uint32_t> midx(std::max<size_t>(off[Mt], 1), 0);
    emu_launch(EmuDim3((Mt + 3) / 4), EmuDim3(128), [&] { fill_missing_kernel(plink, Mt, L, off.data(), midx.data()); });
    std::vector<double> yv((size_t)L.npad, 0.0);
    std::vector<uint8_t> mask4((size_t)L.col_stride, 0);
    for (int i = 0; i < N; i++)
        if ((mask4_in[i / 4] >> (i % 4)) & 1) { yv[i] = y[i]; mask4[i / 4] |= (uint8_t)(1u << (i % 4)); }
    std::vector<double> mave(Mt), msig(Mt), xtx(Mt);
    const int32_t nn = nonas;
    std::vector<uint32_t> na_off, na_idx;
    emu_na_lists(mask4.data(), L.col_stride, 1, N, na_off, na_idx);
    emu_launch(EmuDim3(std::min(Mt, 3)), EmuDim3(kStatsThreads), [&] { stats_kernel(bed.data(), Mt, L, mask4.data(), off.data(), midx.data(), &nn, na_off.data(), na_idx.data(), 1, mave.data(), msig.data(), xtx.data()); });
    std::vector<int32_t> cols(Mt);
    emu_launch(EmuDim3((Mt + 255) / 256), EmuDim3(256), [&] { iota_kernel(cols.data(), Mt); });

    int maxlen = 0;
    auto block = [&](int r, int& S, int& M) { const int size = Mt / R, modu = Mt % R; M = size + (r < modu ? 1 : 0); S = r * size + std::min(r, modu); };
    for (int r = 0; r < R; r++) { int S, M; block(r, S, M); maxlen = std::max(maxlen, M); }
    std::vector<double> part((size_t)gvalue_chunks_of(maxlen) * L.npad), g((size_t)L.npad, 0.0), gk((size_t)L.npad), yk((size_t)L.npad);
    std::vector<double> partial((size_t)vchunk * nsm), spart(nsm), sumsq(1);
    auto gvalues = [&](int b0, int b1, double* add) {
        const int nmark = b1 - b0;
        const GvPlan pl = gvalue_plan(L, nmark);
        if (nmark > 0) {
            emu_launch(EmuDim3(pl.word_blocks, pl.nchunk), EmuDim3(kGvThreads), [&] {
                gvalue_partial_kernel(bed.data(), L.col_stride, pl.nwords, b0, b1, pl.chunk_len, mave.data(), msig.data(), beta_mean, keep, part.data(), L.npad);
            });
            emu_launch(EmuDim3(pl.nchunk), EmuDim3(256), [&] {
                gvalue_missing_kernel(off.data(), midx.data(), b0, b1, pl.chunk_len, mave.data(), msig.data(), beta_mean, keep, part.data(), L.npad);
            });
        }
        emu_launch(EmuDim3((unsigned)((L.npad + 255) / 256)), EmuDim3(256), [&] {
            gvalue_reduce_kernel(part.data(), nmark > 0 ? pl.nchunk : 0, L.npad, L.N, mask4.data(), gk.data(), add);
        });
    };
    for (int r = 0; r < R; r++) { int S, M; block(r, S, M); gvalues(S, S + M, g.data()); }          // pass 1
    memcpy(g_out, g.data(), sizeof(double) * N);
    for (int r = 0; r < R; r++) {                                                                    // pass 2
        int S, M; block(r, S, M);
        if (M == 0) continue;
        gvalues(S, S + M, nullptr);
        emu_launch(EmuDim3((unsigned)((L.npad + 255) / 256)), EmuDim3(256), [&] { predict_residual_kernel(yv.data(), g.data(), gk.data(), L.npad, L.N, yk.data()); });
        emu_launch(EmuDim3(1), EmuDim3(1024), [&] { eps_sumsq_kernel(yk.data(), L.npad, L.N, sumsq.data()); });
        for (int done = 0; done < M; done += vchunk) {
            const int V = std::min(vchunk, M - done);
            StepParams q{};
            q.bed = bed.data(); q.col_stride = L.col_stride; q.nrows = L.nrows; q.cols = cols.data() + S + done; q.V = V; q.eps = yk.data();
            q.npad = L.npad; q.Ttot = 1; q.t0 = 0; q.rows_per_pass = rpp; q.npass = npass; q.partial = partial.data(); q.spart = spart.data();
            q.mask4 = mask4.data(); q.pV = 1; q.err = &err; q.pf = 1;
            emu_launch(EmuDim3(nsm), EmuDim3(kStepThreads), [&] { step_kernel<1>(q); });
            emu_launch(EmuDim3((V + 3) / 4), EmuDim3(128), [&] {
                predict_finish_kernel(cols.data() + S + done, V, nsm, partial.data(), xtx.data(), sumsq.data(), nonas, keep, beta, tdist, se, pval);
            });
        }
    }
    return err;
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    text = open(os.path.join(CSRC, "kernels.cu")).read()
    names = ("helpers", "ingest", "stats", "eps", "step")
    body, _ = asm_to_host.rewrite("".join(text[text.index(f"// [{n}-begin]"):text.index(f"// [{n}-end]")] for n in names))
    ptext = open(os.path.join(CSRC, "predict.cu")).read()
    body += ptext[ptext.index("// [kernels-begin]"):ptext.index("// [kernels-end]")]
    body = body.replace("#pragma unroll\n", "")
    body = body.replace("extern __shared__ __align__(16) uint8_t smem_raw[];", "uint8_t* smem_raw = emu_smem_storage + 16;")
    d = tmp_path_factory.mktemp("emu_predict_chain")
    cpp = d / "predict_chain_emu.cpp"
    cpp.write_text('#include "cuda_emu.h"\n#include "kernels.cuh"\nnamespace gmrm {\n' + body + "\n}\n" + TAIL)
    so = d / "libpredict_chain_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                    "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-unused-function", "-I", os.path.join(EMU, "fake_cuda"), "-I", EMU,
                    "-I", CSRC, str(cpp), "-o", str(so)], check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("N,M,R,nsm,na,miss", [(203, 40, 1, 1, 0.0, 0.0), (780, 60, 3, 2, 0.03, 0.02), (260, 12, 12, 1, 0.0, 0.03)])
def test_emulated_predict_sequence_matches_oracle(emu, oracle, tmp_path, N, M, R, nsm, na, miss):
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=1, n_groups=1, na_rate=na, missing_rate=miss, seed=N % 89)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    rng = np.random.default_rng(M)
    hist = rng.normal(0, 0.02, size=(4, M)) * (rng.random((4, M)) < 0.3)
    keep = (rng.random(M) > 0.05).astype(np.uint8)
    nonas = int(inp["nonas"][0])
    mave, msig = oracle.marker_stats(inp["bed"], N, inp["mask4"][0], nonas)
    want = oracle.predict(inp["bed"], inp["mask4"][0], nonas, inp["eps0"][0], mave, msig, hist, N=N, R=R, keep=keep)
    maxlen = -(-M // R)
    vchunk = max(1, min(maxlen, 1024))                                   # as gmrm_predict chooses it
    plan = api.step_plan(N, nsm, vchunk, 1, want_ranges=False)
    bed = np.ascontiguousarray(inp["bed"], dtype=np.uint8)
    y = np.ascontiguousarray(inp["eps0"][0][:N]); mask4 = np.ascontiguousarray(inp["mask4"][0], dtype=np.uint8)
    bmean = np.ascontiguousarray(hist.mean(axis=0))
    g = np.zeros(N)
    out = {n: np.full(M, 123.0) for n in ("beta", "tdist", "se", "pval")}
    rc = emu.emu_predict(p(bed), N, nsm, M, R, p(y), p(mask4), nonas, p(bmean), p(keep), vchunk, plan["rows_per_pass"], plan["npass"],
                         p(g), p(out["beta"]), p(out["tdist"]), p(out["se"]), p(out["pval"]))
    assert rc == 0
    assert np.abs(g - want["g"][:N]).max() <= 1e-11 * max(np.abs(want["g"]).max(), 1e-300)
    kept = keep != 0
    for n in out:
        assert np.all(np.isnan(out[n][~kept])), n
        np.testing.assert_allclose(out[n][kept], want[n][kept], rtol=1e-9, atol=1e-13, err_msg=n)
