"""Regenerates tests/golden/ from the reference itself (run in the build container, where
/root/reference exists; the GPU box only reads the committed files).

  g1/          N=203 (N%4 != 0), M=400, 2 traits with NAs, missing genotypes, 2 groups x 4 mixtures,
               5 iterations, seed 171014: inputs, the reference's variate logs for 1 and 3 ranks
               (oracle/ref_shim) and the reference's own .bet/.cpn/.csv outputs.
               plus the reference's --predict outputs (.mlma) computed from those .bet histories with 1 and 3
               ranks, against syn.bim / ref.bim (ref.bim: permuted ids, two of them absent).
  lut_ref.bin  dotp_lut_a, dotp_lut_b (1024 doubles each) and na_lut (64) exactly as shipped in the
               reference's src/dotp_lut.hpp and src/na_lut.hpp.
"""
import os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gmrm_b200 import synth            # noqa: E402
from oracle import oracle_py as O      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)
    assert O.have_reference(), "needs oracle/_ref/gmrm_ref (i.e. /root/reference)"
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "lut_ref.bin"), os.path.join(HERE, "lut_ref.bin"))
    g1 = os.path.join(HERE, "g1")
    shutil.rmtree(g1, ignore_errors=True)
    d = synth.write_dataset(g1, N=203, M=400, n_traits=2, n_groups=2, na_rate=0.02, missing_rate=0.01, seed=3)
    p = d["paths"]
    for R in (1, 3):
        out = os.path.join(g1, f"out{R}")
        O.run_reference(g1, p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], out, iterations=5, seed=171014,
                        nranks=R, log_dir=os.path.join(g1, f"log{R}"))
    add_predict(g1, p)
    print("golden fixtures written under", HERE)


def add_predict(g1, p, M=400):
    """Bayes::predict (bayes.cpp:14-284) on the histories written above."""
    bim, ref = os.path.join(g1, "syn.bim"), os.path.join(g1, "ref.bim")
    with open(bim, "w") as f:
        for i in range(M):
            f.write(f"1 rs{i} 0 {1000 + i} A G\n")
    with open(ref, "w") as f:                      # 3 is coprime with 400: a permutation; rs21 and rs222 are dropped
        for i in range(M):
            j = (i * 3) % M
            f.write(f"1 {'rs' + str(j) if j not in (21, 222) else 'absent' + str(j)} 0 {1000 + i} A G\n")
    for R in (1, 3):
        O.run_reference(g1, p["bed"], p["dim"], p["phen"], p["gri"], p["grm"], os.path.join(g1, f"out{R}"), iterations=5,
                        seed=171014, nranks=R, extra=("--predict", "--bim-file", bim, "--ref-bim-file", ref))


if __name__ == "__main__":
    main()
