"""CPU: the marker-statistics kernel (stats_kernel in gmrm_b200/csrc/kernels.cu: mave, 1/sd and the sum of squares the
association pass uses, from dosage counts under the NA mask) run on the host through tests/emu/cuda_emu.h and compared
with the oracle's restatement of PhenMgr::compute_markers_statistics (src/phenotype.cpp:466-556) and with numpy for
sum (a b na)^2 (src/bayes.cpp:190-195).  Test infrastructure only; the GPU parity tests remain the judge of the CUDA build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from gmrm_b200 import synth
from test_predict_kernels_emulated import p, to_device_layout

SRC = os.path.join(ROOT, "gmrm_b200", "csrc", "kernels.cu")
EMU = os.path.join(ROOT, "tests", "emu")

TAIL = r'''
extern "C" void emu_stats(const uint8_t* bed, int nmark, int N, int nsm, const uint8_t* mask4, const uint32_t* miss_off,
                          const uint32_t* miss_idx, const int32_t* nonas, const uint32_t* na_off, const uint32_t* na_idx, int T,
                          double* mave, double* msig, double* xtx, int ctas) {
    using namespace gmrm;
    const Layout L = make_layout(N, nsm);
    emu_launch(EmuDim3(ctas), EmuDim3(kStatsThreads), [&] { stats_kernel(bed, nmark, L, mask4, miss_off, miss_idx, nonas, na_off, na_idx, T, mave, msig, xtx); });
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    text = open(SRC).read()
    a, b = text.index("// [stats-begin]"), text.index("// [stats-end]")
    body = text[a:b].replace("#pragma unroll\n", "")
    d = tmp_path_factory.mktemp("emu_stats")
    cpp = d / "stats_emu.cpp"
    cpp.write_text('#include "cuda_emu.h"\n#include "layout.h"\nnamespace gmrm {\n' + body + "\n}\n" + TAIL)
    so = d / "libstats_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-I", EMU,
                    "-I", os.path.join(ROOT, "gmrm_b200", "csrc"), str(cpp), "-o", str(so)], check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("N,M,T,na,miss", [(203, 40, 2, 0.05, 0.03), (1025, 25, 1, 0.0, 0.0), (4100, 12, 3, 0.02, 0.01)])
def test_emulated_stats_kernel_matches_oracle(emu, oracle, tmp_path, N, M, T, na, miss):
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=T, n_groups=1, na_rate=na, missing_rate=miss, seed=N)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    tri, miss_off, miss_idx, nrows = to_device_layout(inp["bed"], N, 1)
    mask4 = np.zeros((T, nrows * 64), dtype=np.uint8)
    mask4[:, : inp["mask4"].shape[1]] = inp["mask4"]
    nonas = np.ascontiguousarray(inp["nonas"], dtype=np.int32)
    mave = np.full((T, M), np.nan); msig = np.full((T, M), np.nan); xtx = np.full((T, M), np.nan)
    obs_all = ((inp["mask4"][:, :, None] >> np.arange(4)) & 1).reshape(T, -1)[:, :N].astype(bool)
    nas = [np.flatnonzero(~obs_all[t]).astype(np.uint32) for t in range(T)]              # individuals without a phenotype, per trait
    na_off = np.concatenate([[0], np.cumsum([x.size for x in nas])]).astype(np.uint32)
    na_idx = np.concatenate(nas + [np.zeros(1, np.uint32)]).astype(np.uint32)
    emu.emu_stats(p(tri), M, N, 1, p(mask4), p(miss_off), p(miss_idx), p(nonas), p(na_off), p(na_idx), T, p(mave), p(msig), p(xtx),
                  3)                                                                       # 3 persistent CTAs share the M markers
    codes = (inp["bed"][:, :, None] >> (2 * np.arange(4))) & 3
    a = np.where(codes == 0, 2.0, np.where(codes == 2, 1.0, 0.0)).reshape(M, -1)[:, :N]
    for t in range(T):
        mave_o, msig_o = oracle.marker_stats(inp["bed"], N, inp["mask4"][t], int(inp["nonas"][t]))
        np.testing.assert_allclose(mave[t], mave_o, rtol=1e-13)
        np.testing.assert_allclose(msig[t], msig_o, rtol=1e-12)
        obs = ((inp["mask4"][t][:, None] >> np.arange(4)) & 1).reshape(-1)[:N].astype(bool)
        assert np.array_equal(xtx[t], ((a * a) * obs).sum(axis=1))          # integer-valued: exact
