"""Compares gpurun_out/pred{1,3}/*.mlma (written by the CLI on the GPU box) with the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle_py as O
D = "gpurun_in/pred"
inp = O.load_inputs(D + "/syn.bed", D + "/syn.dim", [D + "/syn_t0.phen", D + "/syn_t1.phen"], D + "/syn.gri", D + "/syn.grm")
hists, keep = np.load(D + "/hists.npy"), np.load(D + "/keep.npy")
ok = True
for R in (1, 3):
    for t in range(2):
        path = f"gpurun_out/pred{R}/syn_t{t}.mlma"
        if not os.path.exists(path):
            print("missing", path); ok = False; continue
        rows = O.read_mlma(path)
        mave, msig = O.marker_stats(inp["bed"], inp["N"], inp["mask4"][t], int(inp["nonas"][t]))
        want = O.predict(inp["bed"], inp["mask4"][t], int(inp["nonas"][t]), inp["eps0"][t], mave, msig, hists[t], N=inp["N"], R=R, keep=keep)
        idx = np.array([r[1] for r in rows])
        good = np.array_equal(idx, np.flatnonzero(keep)) and os.path.getsize(path) == 123 * int(keep.sum())
        errs = {}
        for c, name in ((3, "beta"), (4, "tdist"), (5, "se"), (6, "pval")):
            got = np.array([r[c] for r in rows])
            errs[name] = float(np.max(np.abs(got - want[name][idx]) / np.maximum(np.abs(want[name][idx]), 1e-3))) if good else float("nan")
        print(f"R={R} t={t} rows={len(rows)} order_ok={good} max rel err {errs}")
        ok &= good and max(errs.values()) < 1e-9
print("PREDICT_HW_OK" if ok else "PREDICT_HW_MISMATCH")
