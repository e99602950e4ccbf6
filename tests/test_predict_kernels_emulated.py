"""CPU: the kernels of the association pass (gmrm_b200/csrc/predict.cu) executed on the host through a small CUDA
emulator (tests/emu/cuda_emu.h: CTAs in turn, one std::thread per CUDA thread, barriers for __syncthreads and warp
shuffles) and compared with the oracle's restatement of Bayes::predict (src/bayes.cpp:87-214).

This checks the kernels' indexing, chunking, missing-genotype correction and statistics without a GPU -- the text
between the [kernels-begin] / [kernels-end] markers of predict.cu is compiled verbatim.  It is test infrastructure: the
product has no CPU path, and the GPU parity tests (tests/test_gpu_predict.py) remain the judge of the CUDA build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from gmrm_b200 import synth

SRC = os.path.join(ROOT, "gmrm_b200", "csrc", "predict.cu")
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    text = open(SRC).read()
    a, b = text.index("// [kernels-begin]"), text.index("// [kernels-end]")
    body = text[a:b].replace("#pragma unroll\n", "")
    d = tmp_path_factory.mktemp("emu")
    cpp = d / "predict_emu.cpp"
    cpp.write_text('#include "cuda_emu.h"\n#include "layout.h"\nnamespace gmrm {\nnamespace {\n' + body + "\n}\n}\n"
                   + open(os.path.join(EMU, "predict_emu_tail.inc")).read())
    so = d / "libpredict_emu.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
                    "-I", EMU, "-I", os.path.join(ROOT, "gmrm_b200", "csrc"), str(cpp), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def to_device_layout(bed, N, nsm):
    """layout.h: PLINK bytes -> base-3 quad bytes (missing stored as dosage 0) in columns padded to rows of 64 bytes,
    plus the per-marker CSR list of missing individuals."""
    M, mbytes = bed.shape
    nrows = -(-mbytes // 64)
    codes = (bed[:, :, None] >> (2 * np.arange(4))) & 3                        # [M][mbytes][4]
    dos = np.where(codes == 0, 2, np.where(codes == 2, 1, 0)).astype(np.uint8)
    tri = (dos * np.array([1, 3, 9, 27], dtype=np.uint8)).sum(axis=2).astype(np.uint8)
    out = np.zeros((M, nrows * 64), dtype=np.uint8)
    out[:, :mbytes] = tri
    miss = codes == 1
    miss_off = np.zeros(M + 1, dtype=np.uint32)
    idx = []
    for m in range(M):
        ind = np.flatnonzero(miss[m].reshape(-1))
        ind = ind[ind < N]
        idx.append(ind)
        miss_off[m + 1] = miss_off[m] + ind.size
    miss_idx = np.concatenate(idx).astype(np.uint32) if idx else np.zeros(0, np.uint32)
    if miss_idx.size == 0:
        miss_idx = np.zeros(1, np.uint32)
    return out, miss_off, miss_idx, nrows


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("N,M,R,nsm,na,miss", [(203, 70, 1, 1, 0.02, 0.02), (1030, 200, 3, 2, 0.03, 0.01), (515, 2200, 2, 1, 0.0, 0.005),
                                               (64, 33, 33, 1, 0.0, 0.0)])
def test_emulated_kernels_match_oracle(emu, oracle, tmp_path, N, M, R, nsm, na, miss):
    d = synth.write_dataset(str(tmp_path), N=N, M=M, n_traits=1, n_groups=1, na_rate=na, missing_rate=miss, seed=N + M)
    pp = d["paths"]
    inp = oracle.load_inputs(pp["bed"], pp["dim"], pp["phen"], pp["gri"], pp["grm"])
    rng = np.random.default_rng(R)
    hist = rng.normal(0, 0.05, size=(3, M)) * (rng.random((3, M)) < 0.4)
    keep = (rng.random(M) > 0.06).astype(np.uint8)
    nonas = int(inp["nonas"][0])
    mave, msig = oracle.marker_stats(inp["bed"], N, inp["mask4"][0], nonas)
    want = oracle.predict(inp["bed"], inp["mask4"][0], nonas, inp["eps0"][0], mave, msig, hist, N=N, R=R, keep=keep)

    tri, miss_off, miss_idx, nrows = to_device_layout(inp["bed"], N, nsm)
    npad = nrows * 256
    mask4 = np.zeros(nrows * 64, dtype=np.uint8)
    mask4[: inp["mask4"][0].size] = inp["mask4"][0]
    y = np.zeros(npad)
    y[:N] = inp["eps0"][0][:N]
    bmean = np.ascontiguousarray(hist.mean(axis=0))
    cols = np.full(M, -1, dtype=np.int32)
    emu.emu_iota(p(cols), M)
    assert np.array_equal(cols, np.arange(M))

    g = np.zeros(npad)
    gk = np.full(npad, np.nan)
    blocks = [oracle.block_of_markers(M, R, r)[:2] for r in range(R)]
    for S, Mr in blocks:                                                    # pass 1 of gmrm_predict
        emu.emu_gvalues(p(tri), N, nsm, p(miss_off), p(miss_idx), S, S + Mr, p(mave), p(msig), p(bmean), p(keep), p(mask4), p(gk), p(g))
    scale = max(np.abs(want["g"]).max(), 1e-300)
    assert np.abs(g[:N] - want["g"][:N]).max() <= 1e-11 * scale
    assert not g[N:].any()

    # dosages with missing as 0 and the NA mask, for the marker sums the step kernel delivers on the GPU
    codes = (inp["bed"][:, :, None] >> (2 * np.arange(4))) & 3
    a = np.where(codes == 0, 2.0, np.where(codes == 2, 1.0, 0.0)).reshape(M, -1)[:, :N]
    obs = ((inp["mask4"][0][:, None] >> np.arange(4)) & 1).reshape(-1)[:N].astype(bool)
    xtx = ((a * a) * obs).sum(axis=1)
    out = {n: np.full(M, 123.0) for n in ("beta", "tdist", "se", "pval")}
    nsm_part = 5                                                            # the step kernel's per-CTA partial sums, faked by a 5-way split
    for S, Mr in blocks:                                                    # pass 2
        emu.emu_gvalues(p(tri), N, nsm, p(miss_off), p(miss_idx), S, S + Mr, p(mave), p(msig), p(bmean), p(keep), p(mask4), p(gk), None)
        yk = np.full(npad, np.nan)
        emu.emu_residual(p(y), p(g), p(gk), N, nsm, p(yk))
        assert not yk[N:].any()
        sumsq = np.array([float(np.dot(yk[:N], yk[:N]))])
        xty = a[S:S + Mr] @ yk[:N]
        partial = np.zeros((Mr, nsm_part))
        partial[:, 0] = 0.25 * xty; partial[:, 3] = 0.5 * xty; partial[:, 4] = 0.25 * xty
        emu.emu_finish(p(cols[S:]), Mr, nsm_part, p(partial), p(xtx), p(sumsq), nonas, p(keep),
                       p(out["beta"]), p(out["tdist"]), p(out["se"]), p(out["pval"]))
    kept = keep != 0
    for n in out:
        assert np.all(np.isnan(out[n][~kept])), n
        np.testing.assert_allclose(out[n][kept], want[n][kept], rtol=1e-9, atol=1e-13, err_msg=n)
