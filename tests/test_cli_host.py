"""Host side (gmrm_b200/gmrm_b200_cli): the reference's command line and file formats.
CPU tests cover the parser and the input readers (--check-inputs does no GPU work); the GPU test runs the
executable end to end and compares its .bet / .cpn / .csv files with the oracle following the same Philox streams."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from gmrm_b200 import synth

CLI = os.path.join(ROOT, "gmrm_b200", "gmrm_b200_cli")


def run(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=600, **kw)


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    if not os.path.exists(CLI):
        pytest.skip("gmrm_b200_cli not built")
    tmp = tmp_path_factory.mktemp("cli")
    d = synth.write_dataset(str(tmp), N=1003, M=300, n_traits=2, n_groups=2, na_rate=0.01, missing_rate=0.005, seed=3)
    d["tmp"] = str(tmp)
    return d


def base_args(d, out):
    p = d["paths"]
    return ["--bed-file", p["bed"], "--dim-file", p["dim"], "--phen-files", ",".join(p["phen"]), "--group-index-file", p["gri"],
            "--group-mixture-file", p["grm"], "--out-dir", out]


def test_cli_echoes_options_and_counts_nas(data, oracle):
    r = run(base_args(data, os.path.join(data["tmp"], "o1")) + ["--iterations", "3", "--seed", "5", "--check-inputs"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ardyh command line options:" in r.stdout and "--iterations 3" in r.stdout      # options.cpp:22,158-159
    assert "Setting last 1 bits to NAs" in r.stdout                                        # phenotype.cpp:633-638 (N = 1003)
    inp = oracle.load_inputs(data["paths"]["bed"], data["paths"]["dim"], data["paths"]["phen"], data["paths"]["gri"], data["paths"]["grm"])
    for t, ph in enumerate(data["paths"]["phen"]):
        assert f"{ph}: {int(inp['nonas'][t])} observed, {1003 - int(inp['nonas'][t])} NA" in r.stdout


def test_cli_readers_match_oracle(data, oracle, tmp_path):
    """.phen / .gri / .grm readers of the executable against the oracle's restatement of Phenotype::read_file
    (phenotype.cpp:587-673), read_group_index_file (bayes.cpp:830-853) and read_group_mixture_file (options.cpp:222-286):
    centred-scaled phenotypes to 1e-15, NA masks / groups / mixtures bit-exact."""
    r = run(base_args(data, str(tmp_path / "o")) + ["--check-inputs", "--dump-inputs", str(tmp_path)])
    assert r.returncode == 0, r.stdout + r.stderr
    p = data["paths"]
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    for t in range(2):
        eps = np.fromfile(tmp_path / f"eps{t}.f64")
        mask = np.fromfile(tmp_path / f"mask{t}.u8", dtype=np.uint8)
        assert np.array_equal(mask, inp["mask4"][t])
        np.testing.assert_allclose(eps[: inp["N"]], inp["eps0"][t][: inp["N"]], rtol=0, atol=1e-15)
        assert (eps[inp["N"]:] == 0).all()
    assert np.array_equal(np.fromfile(tmp_path / "groups.i32", dtype=np.int32), inp["group_index"])
    assert np.array_equal(np.fromfile(tmp_path / "cva.f64").reshape(inp["cva"].shape), inp["cva"])


def test_cli_writers_produce_the_reference_layouts(data, oracle, tmp_path):
    """.csv / .bet / .cpn writers (xfiles.cpp:17-45, xfiles.hpp:24-37) driven as the iteration loop drives them from two
    ranks with thin rate 2; read back with the parsers that tests/test_oracle_golden.py pins on the reference's own files."""
    out = tmp_path / "o"
    r = run(base_args(data, str(out)) + ["--check-inputs", "--selftest-outputs"])
    assert r.returncode == 0, r.stdout + r.stderr
    Mt, G, K = 300, 2, 4
    its, bet = oracle.read_bet(str(out / "selftest.bet"))
    its_c, cpn = oracle.read_cpn(str(out / "selftest.cpn"))
    assert list(its) == [2, 4] and list(its_c) == [2, 4]
    assert os.path.getsize(out / "selftest.bet") == 4 + 2 * (4 + 8 * Mt) and os.path.getsize(out / "selftest.cpn") == 4 + 2 * (4 + 4 * Mt)
    j = np.arange(Mt)
    for n, it in enumerate((2, 4)):
        np.testing.assert_array_equal(bet[n], 1e-3 * j - 0.5 * it)
        np.testing.assert_array_equal(cpn[n], (j + it) % K)
    lines = open(out / "selftest.csv").read().splitlines()
    assert len(lines) == 2 and len(set(len(x) for x in lines)) == 1                        # constant line length (xfiles.cpp:45)
    for n, it in enumerate((2, 4)):
        sg = [0.1 * (g + 1) + 0.001 * it for g in range(G)]
        want = "%5d, %4d" % (it, G) + "".join(", %20.15f" % v for v in sg)
        sige = 0.5 + 0.01 * it
        want += ", %20.15f, %20.15f, %7d, %4d, %2d" % (sige, sum(sg) / (sige + sum(sg)), 1234 + it, G, K)
        want += "".join(", %20.15f" % ((i + 1.0) / (G * K * 10.0) + 1e-4 * it) for i in range(G * K))
        assert lines[n] == want
    rows = oracle.read_csv(str(out / "selftest.csv"))
    assert [r_["it"] for r_ in rows] == [2, 4] and rows[1]["m0_sum"] == 1238


@pytest.mark.parametrize("args,msg", [(["--bogus", "1"], 'option "--bogus" unknown'),                       # options.cpp:152-155
                                      (["--iterations"], "missing argument for last option"),               # options.cpp:169-172
                                      (["--iterations", "0"], "strictly positive"),
                                      ([], "no bed file provided")])
def test_cli_rejects_bad_command_lines(data, args, msg):
    r = run(args)
    assert r.returncode == 1
    assert msg in r.stdout


def test_cli_needs_both_group_files(data):
    p = data["paths"]
    r = run(["--bed-file", p["bed"], "--dim-file", p["dim"], "--phen-files", p["phen"][0], "--group-index-file", p["gri"]])
    assert r.returncode == 1 and "BOTH --group-index-file and --group-mixture-file" in r.stdout   # options.cpp:198-203


def test_cli_rejects_unsorted_mixtures(data, tmp_path):
    bad = tmp_path / "bad.grm"
    bad.write_text("0.0 0.01 0.001 0.1\n0.0 0.001 0.01 0.1\n")
    args = base_args(data, str(tmp_path / "o"))
    args[args.index("--group-mixture-file") + 1] = str(bad)
    r = run(args + ["--check-inputs"])
    assert r.returncode == 1 and "ascending order" in r.stdout                              # options.cpp:278-281


@pytest.mark.gpu
@pytest.mark.parametrize("vranks,thin", [(1, 1), (16, 2)])
def test_cli_outputs_match_oracle(data, oracle, vranks, thin):
    out = os.path.join(data["tmp"], f"out_{vranks}_{thin}")
    iters, seed = 6, 4242
    # the second case loads the .bed in many small chunks: both pinned buffers of the pipelined reader take turns
    env = dict(os.environ, GMRM_CLI_CHUNK_KB="8") if vranks == 16 else None
    r = run(base_args(data, out) + ["--iterations", str(iters), "--seed", str(seed), "--vranks", str(vranks), "--output-thin-rate", str(thin),
                                     "--burn-in", "2"], env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("total proc time") == iters                                      # bayes.cpp:655
    p = data["paths"]
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=inp["N"], R=vranks,
                       iterations=iters, rng_mode=1, seed=seed)
    saved = [i for i in range(1, iters + 1) if i % thin == 0]
    for t in range(2):
        stem = os.path.splitext(os.path.basename(p["phen"][t]))[0]
        its, bet = oracle.read_bet(os.path.join(out, stem + ".bet"))
        its_c, cpn = oracle.read_cpn(os.path.join(out, stem + ".cpn"))
        assert list(its) == saved and list(its_c) == saved
        for n, it in enumerate(saved):
            assert np.array_equal(cpn[n], res["comp"][it - 1][t])
            np.testing.assert_allclose(bet[n], res["betas"][it - 1][t], rtol=1e-8, atol=1e-13)
        csv = oracle.read_csv(os.path.join(out, stem + ".csv"))
        assert len(csv) == len(saved)
        for n, it in enumerate(saved):
            assert int(csv[n]["it"]) == it
            np.testing.assert_allclose(csv[n]["sigmag"], res["sigmag"][it - 1][t], rtol=1e-8)
            np.testing.assert_allclose(csv[n]["sigmae"], res["sigmae"][it - 1][t], rtol=1e-8)
        mb = np.fromfile(os.path.join(out, stem + ".mbet"))
        np.testing.assert_allclose(mb, np.mean([res["betas"][i][t] for i in range(2, iters)], axis=0), rtol=1e-8, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("sync_rate", [1, 3])
def test_cli_two_gpus_match_oracle(data, oracle, sync_rate):
    """--gpus 2: two host threads in one process, one engine each, peers' buffers passed as plain pointers (sync rate 1: list
    exchange over peer memory; sync rate 3: all-reduced residual deltas) against the oracle with the same 16 virtual ranks."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = os.path.join(data["tmp"], f"out_2gpu_{sync_rate}")
    iters, seed, vranks = 4, 77, 16
    r = run(base_args(data, out) + ["--iterations", str(iters), "--seed", str(seed), "--vranks", str(vranks), "--gpus", "2",
                                     "--sync-rate", str(sync_rate)])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    p = data["paths"]
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    res = oracle.gibbs(inp["bed"], inp["eps0"], inp["mask4"], inp["nonas"], inp["group_index"], inp["cva"], N=inp["N"], R=vranks,
                       nrep=1 if sync_rate == 1 else 2, iterations=iters, rng_mode=1, seed=seed, sync_rate=sync_rate)
    for t in range(2):
        stem = os.path.splitext(os.path.basename(p["phen"][t]))[0]
        its, bet = oracle.read_bet(os.path.join(out, stem + ".bet"))
        _, cpn = oracle.read_cpn(os.path.join(out, stem + ".cpn"))
        assert list(its) == list(range(1, iters + 1))
        for i in range(iters):
            assert np.array_equal(cpn[i], res["comp"][i][t])
            np.testing.assert_allclose(bet[i], res["betas"][i][t], rtol=1e-8, atol=1e-13)
        csv = oracle.read_csv(os.path.join(out, stem + ".csv"))
        for i in range(iters):
            np.testing.assert_allclose(csv[i]["sigmag"], res["sigmag"][i][t], rtol=1e-8)
            np.testing.assert_allclose(csv[i]["sigmae"], res["sigmae"][i][t], rtol=1e-8)


# ------------------------------------------------------------------ --predict (Bayes::predict, bayes.cpp:14-284)
def write_predict_inputs(d, out, M, niter=3, seed=5):
    """A .bim pair (reference reversed, every 50th id replaced) and a .bet history per trait, as the Gibbs mode writes it."""
    os.makedirs(out, exist_ok=True)
    bim, ref = os.path.join(d["tmp"], "p.bim"), os.path.join(d["tmp"], "pref.bim")
    with open(bim, "w") as f:
        f.writelines(f"1 rs{i} 0 {i} A G\n" for i in range(M))
    with open(ref, "w") as f:
        f.writelines(f"1 {'rs' if i % 50 else 'gone'}{i} 0 {i} A G\n" for i in reversed(range(M)))
    rng = np.random.default_rng(seed)
    hists = []
    for ph in d["paths"]["phen"]:
        stem = os.path.splitext(os.path.basename(ph))[0]
        h = rng.normal(0, 0.02, size=(niter, M)) * (rng.random((niter, M)) < 0.3)
        with open(os.path.join(out, stem + ".bet"), "wb") as f:                    # xfiles.hpp:24-37
            f.write(np.uint32(M).tobytes())
            for i in range(niter):
                f.write(np.uint32(i + 1).tobytes())
                f.write(h[i].tobytes())
        hists.append(h)
    keep = np.array([i % 50 != 0 for i in range(M)], dtype=np.uint8)
    return bim, ref, hists, keep


def predict_args(d, out, bim, ref):
    p = d["paths"]
    return ["--bed-file", p["bed"], "--dim-file", p["dim"], "--phen-files", ",".join(p["phen"]), "--out-dir", out,
            "--predict", "--bim-file", bim, "--ref-bim-file", ref]


def test_cli_predict_needs_both_bim_files(data):
    p = data["paths"]
    base = ["--bed-file", p["bed"], "--dim-file", p["dim"], "--phen-files", p["phen"][0], "--predict"]
    r = run(base)
    assert r.returncode == 1 and "you need to pass a bim file with --bim-file" in r.stdout                 # options.cpp:206-208
    r = run(base + ["--bim-file", "x.bim"])
    assert r.returncode == 1 and "you need to pass a reference bim file with --ref-bim-file" in r.stdout   # options.cpp:210-212


def test_cli_predict_readers(data, tmp_path):
    """.bim cross-reference (bayes.cpp:286-316), .bet mean (38-78) and the .mlma line format (230-236) on the CPU."""
    out = str(tmp_path / "o")
    bim, ref, hists, keep = write_predict_inputs(data, out, 300)
    r = run(predict_args(data, out, bim, ref) + ["--check-inputs", "--selftest-predict"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "INFO   : found 300 ids in bim file" in r.stdout and "INFO   : found 300 ids in reference bim file" in r.stdout
    assert "Number of recorded iterations in .bet file 1: 3" in r.stdout                                    # bayes.cpp:52-53
    assert f"SELFTEST: kept {int(keep.sum())} of 300 markers" in r.stdout
    w = np.arange(1, 301)
    for t in range(2):
        line = [l for l in r.stdout.splitlines() if l.startswith(f"SELFTEST: trait {t} weighted beta mean")][0]
        assert abs(float(line.split()[-1]) - float((hists[t].mean(axis=0) * w).sum())) < 1e-12
    want = "%20s %8d %8d %20.15f %20.15f %20.15f %20.15f" % ("rs299", 299, 0, 0.25, -1.5, 0.125, 0.0625)
    assert "SELFTEST: " + want in r.stdout and len(want) == 122


def test_cli_predict_rejects_a_history_of_another_size(data, tmp_path):
    out = str(tmp_path / "o")
    bim, ref, _, _ = write_predict_inputs(data, out, 300)
    with open(ref, "a") as f:
        f.write("1 extra 0 1 A G\n")                       # 301 reference ids, 300 markers in the .bet
    r = run(predict_args(data, out, bim, ref) + ["--check-inputs"])
    assert r.returncode == 1 and "Mismatch between expected and Mtot read from .bet file" in r.stdout       # bayes.cpp:45-48


@pytest.mark.gpu
@pytest.mark.parametrize("vranks", [1, 3])
def test_cli_predict_matches_oracle(data, oracle, vranks):
    out = os.path.join(data["tmp"], f"pred_{vranks}")
    bim, ref, hists, keep = write_predict_inputs(data, out, 300)
    r = run(predict_args(data, out, bim, ref) + ["--vranks", str(vranks)])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("excluded -- no match") == 2 * int((keep == 0).sum())                             # bayes.cpp:227, per trait
    p = data["paths"]
    inp = oracle.load_inputs(p["bed"], p["dim"], p["phen"], p["gri"], p["grm"])
    for t in range(2):
        stem = os.path.splitext(os.path.basename(p["phen"][t]))[0]
        rows = oracle.read_mlma(os.path.join(out, stem + ".mlma"))
        assert os.path.getsize(os.path.join(out, stem + ".mlma")) == 123 * int(keep.sum())
        mave, msig = oracle.marker_stats(inp["bed"], inp["N"], inp["mask4"][t], int(inp["nonas"][t]))
        want = oracle.predict(inp["bed"], inp["mask4"][t], int(inp["nonas"][t]), inp["eps0"][t], mave, msig, hists[t], N=inp["N"],
                              R=vranks, keep=keep)
        idx = np.array([r_[1] for r_ in rows])
        assert np.array_equal(idx, np.flatnonzero(keep)) and [r_[2] for r_ in rows] == [299 - i for i in idx]
        for c, name in ((3, "beta"), (4, "tdist"), (5, "se"), (6, "pval")):
            np.testing.assert_allclose([r_[c] for r_ in rows], want[name][idx], rtol=1e-9, atol=2e-15, err_msg=name)


def test_cli_trunc_markers(data, tmp_path):
    """--trunc-markers n keeps the first n markers (dimensions.hpp:12-14); the group file may list more (bayes.cpp:844-852)."""
    r = run(base_args(data, str(tmp_path / "o")) + ["--trunc-markers", "120", "--check-inputs", "--dump-inputs", str(tmp_path)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "--trunc-markers 120" in r.stdout and "M = 120 markers" in r.stdout
    groups = np.fromfile(tmp_path / "groups.i32", dtype=np.int32)
    want = np.array([int(l.split()[1]) for l in open(data["paths"]["gri"])], dtype=np.int32)
    assert groups.size >= 120 and np.array_equal(groups[:120], want[:120])
    r = run(base_args(data, str(tmp_path / "o")) + ["--trunc-markers", "100000", "--check-inputs"])   # larger than Mt: no effect
    assert r.returncode == 0 and "M = 300 markers" in r.stdout
