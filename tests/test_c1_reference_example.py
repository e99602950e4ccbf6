"""BASELINE.json configs[0]: the reference's bundled example (example/part1_gcc_mvapich2.sh:15-24) -- test.dim, test.gri, test.grm
and the phenotypes test1, test1_bis (== test1), test1_nas (one NA, line 9) and test2; 2 ranks, seed 171014, 10 iterations.

The text fixtures under tests/golden/c1/ are byte copies of the reference's files (gzip-compressed; test1_bis and test1_nas are
rebuilt from test1 as the reference ships them); example/test.bed is not in the reference checkout (.MISSING_LARGE_BLOBS) and
cannot be regenerated without R and plink, so a stand-in .bed of the same shape is used (Binomial(2, 0.4) genotypes,
data_sim.R:15).  The bundled phenotypes therefore carry no signal about these genotypes; what is tested is the path:

  CPU  * fixtures are byte-identical to /root/reference/example (where that exists)
       * the executable's readers parse the reference's own files exactly like the oracle's restatement of the reference readers
  GPU  * the unmodified reference binary (oracle/_ref, 2 shim ranks) runs the example; its logged variates are replayed through
         gmrm_b200_cli --replay-file; .cpn equal, .bet / .csv to 1e-6 relative (the north-star replay bound)
       * test1 and test1_bis give identical chains; the NA of test1_nas is handled (nonas = 9,999) and its chain differs."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from gmrm_b200 import synth

CLI = os.path.join(ROOT, "gmrm_b200", "gmrm_b200_cli")
C1 = os.path.join(GOLDEN, "c1")
REF_EXAMPLE = "/root/reference/example"
STEMS = ["test1", "test1_bis", "test1_nas", "test2"]          # order of --phen-files in part1_gcc_mvapich2.sh:19


def materialise(dst):
    """The example directory as the reference ships it (minus test.bed), rebuilt from the fixtures."""
    os.makedirs(dst, exist_ok=True)
    for name in ("test.dim", "test.grm"):
        shutil.copy(os.path.join(C1, name), os.path.join(dst, name))
    for name in ("test.gri", "test1.phen", "test2.phen"):
        with gzip.open(os.path.join(C1, name + ".gz"), "rb") as f, open(os.path.join(dst, name), "wb") as g:
            g.write(f.read())
    lines = open(os.path.join(dst, "test1.phen")).read().splitlines(keepends=True)
    open(os.path.join(dst, "test1_bis.phen"), "w").write("".join(lines))
    lines[8] = "9 9 NA\n"                                         # example/test1_nas.phen:9
    open(os.path.join(dst, "test1_nas.phen"), "w").write("".join(lines))
    return dst


def standin_bed(dst, N, M):
    """Stand-in for the missing example/test.bed: rbinom(N*M, 2, 0.4) dosages (data_sim.R:15), AA -> 00, AG -> 10, GG -> 11."""
    d = synth.make_genotypes(N, M, seed=171014, maf_lo=0.4, maf_hi=0.4)
    with open(os.path.join(dst, "test.bed"), "wb") as f:
        f.write(synth.BED_MAGIC)
        f.write(synth.pack_bed(d).tobytes())


def args_for(d, out, extra=()):
    return ["--bed-file", os.path.join(d, "test.bed"), "--dim-file", os.path.join(d, "test.dim"),
            "--phen-files", ",".join(os.path.join(d, s + ".phen") for s in STEMS),
            "--group-index-file", os.path.join(d, "test.gri"), "--group-mixture-file", os.path.join(d, "test.grm"),
            "--shuffle-markers", "1", "--seed", "171014", "--iterations", "10", "--out-dir", out, *extra]


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLE), reason="/root/reference is not on this machine")
def test_fixtures_are_the_references_files(tmp_path):
    d = materialise(str(tmp_path))
    for name in ("test.dim", "test.grm", "test.gri", "test1.phen", "test1_bis.phen", "test1_nas.phen", "test2.phen"):
        assert open(os.path.join(d, name), "rb").read() == open(os.path.join(REF_EXAMPLE, name), "rb").read(), name


def test_cli_readers_on_the_references_files(oracle, tmp_path):
    """Dimensions, the 20,000-line group index, the mixtures and the four phenotypes (one NA) through the executable's readers
    (--check-inputs: no GPU work) against the oracle's restatement of the reference's readers: masks / groups / mixtures
    bit-exact, centred-scaled phenotypes to 1e-15."""
    if not os.path.exists(CLI):
        pytest.skip("gmrm_b200_cli not built")
    d = materialise(str(tmp_path / "ex"))
    N, M = (int(x) for x in open(os.path.join(d, "test.dim")).read().split())
    assert (N, M) == (10000, 20000)
    open(os.path.join(d, "test.bed"), "wb").write(synth.BED_MAGIC)           # --check-inputs never opens the genotypes
    dump = tmp_path / "dump"
    dump.mkdir()
    r = subprocess.run([CLI] + args_for(d, str(tmp_path / "o"), ["--check-inputs", "--dump-inputs", str(dump)]),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "N = 10000 individuals, M = 20000 markers, 4 trait(s), 1 group(s) x 4 mixtures" in r.stdout
    assert f"{os.path.join(d, 'test1_nas.phen')}: 9999 observed, 1 NA" in r.stdout
    assert f"{os.path.join(d, 'test1.phen')}: 10000 observed, 0 NA" in r.stdout
    assert "312 virtual ranks" in r.stdout                                   # default: Mt / 64 (host.hpp default_vranks), not Mt
    eps = []
    for t, s in enumerate(STEMS):
        e, m, nonas, nas = oracle.read_phen(os.path.join(d, s + ".phen"), N)
        got = np.fromfile(dump / f"eps{t}.f64")
        assert np.array_equal(np.fromfile(dump / f"mask{t}.u8", dtype=np.uint8), m)
        np.testing.assert_allclose(got[:N], e[:N], rtol=0, atol=1e-15)
        assert nas == (1 if s == "test1_nas" else 0)
        eps.append(got)
    assert np.array_equal(eps[0], eps[1]) and not np.array_equal(eps[0], eps[2])
    assert eps[2][8] == 0.0                                                   # the NA individual
    gi = np.fromfile(dump / "groups.i32", dtype=np.int32)
    assert gi.size == M and not gi.any()
    assert np.array_equal(np.fromfile(dump / "cva.f64"), np.array([0.0, 1e-4, 1e-3, 1e-2]))


@pytest.mark.gpu
def test_c1_replay_through_the_cli_matches_the_reference(oracle, tmp_path):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/gmrm_ref not present")
    d = materialise(str(tmp_path / "ex"))
    N, M, R, iters = 10000, 20000, 2, 10
    standin_bed(d, N, M)
    phens = [os.path.join(d, s + ".phen") for s in STEMS]
    log, ref_out, our_out = str(tmp_path / "log"), str(tmp_path / "ref"), str(tmp_path / "ours")
    # the reference itself, as example/part1_gcc_mvapich2.sh runs it: 2 ranks, seed 171014, 10 iterations
    oracle.run_reference(d, os.path.join(d, "test.bed"), os.path.join(d, "test.dim"), phens, os.path.join(d, "test.gri"),
                         os.path.join(d, "test.grm"), ref_out, iterations=iters, seed=171014, nranks=R, log_dir=log, timeout=2400)
    inp = oracle.load_inputs(os.path.join(d, "test.bed"), os.path.join(d, "test.dim"), phens, os.path.join(d, "test.gri"),
                             os.path.join(d, "test.grm"))
    assert list(inp["nonas"]) == [10000, 10000, 9999, 10000]
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=N, R=R,
                       iterations=iters, rng_mode=0, replay_dir=log)
    rpl = str(tmp_path / "c1.rpl")
    oracle.write_replay_file(rpl, res)
    r = subprocess.run([CLI] + args_for(d, our_out, ["--replay-file", rpl]), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "replaying the logged variates of 10 iterations of a 2-rank reference run" in r.stdout
    ours = {}
    for s in STEMS:
        its_r, bet_r = oracle.read_bet(os.path.join(ref_out, s + ".bet"))
        its_o, bet_o = oracle.read_bet(os.path.join(our_out, s + ".bet"))
        _, cpn_r = oracle.read_cpn(os.path.join(ref_out, s + ".cpn"))
        _, cpn_o = oracle.read_cpn(os.path.join(our_out, s + ".cpn"))
        assert list(its_o) == list(its_r) == list(range(1, iters + 1))
        assert np.array_equal(cpn_o, cpn_r), s                                 # components: bit-exact
        np.testing.assert_allclose(bet_o, bet_r, rtol=1e-6, atol=1e-12, err_msg=s)
        csv_r, csv_o = oracle.read_csv(os.path.join(ref_out, s + ".csv")), oracle.read_csv(os.path.join(our_out, s + ".csv"))
        assert len(csv_o) == len(csv_r) == iters
        for a, b in zip(csv_o, csv_r):
            assert a["it"] == b["it"] and a["m0_sum"] == b["m0_sum"]
            np.testing.assert_allclose(a["sigmag"], b["sigmag"], rtol=1e-6)
            np.testing.assert_allclose(a["sigmae"], b["sigmae"], rtol=1e-6)
            np.testing.assert_allclose(a["pi"], b["pi"], rtol=1e-6)
        ours[s] = (bet_o, cpn_o)
    # identical phenotypes, identically seeded traits (bayes.cpp:796-803): identical chains; the NA makes a different one
    assert np.array_equal(ours["test1"][0], ours["test1_bis"][0]) and np.array_equal(ours["test1"][1], ours["test1_bis"][1])
    assert not np.array_equal(ours["test1"][0], ours["test1_nas"][0])
