"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_build/liboracle.so (see oracle.h) plus
helpers to run the reference binary oracle/_ref/gmrm_ref and to parse the reference's output
files.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "gmrm_ref")


def build(quiet: bool = True) -> None:
    """make the oracle library (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class OracleCfg(C.Structure):
    _fields_ = [("N", C.c_int), ("Mt", C.c_int), ("T", C.c_int), ("G", C.c_int), ("K", C.c_int),
                ("R", C.c_int), ("nrep", C.c_int), ("iterations", C.c_int), ("shuffle", C.c_int),
                ("rng_mode", C.c_int), ("sync_rate", C.c_int), ("seed", C.c_uint32),
                ("replay_dir", C.c_char_p)]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)


class OracleOut(C.Structure):
    _fields_ = [("betas", _DP), ("comp", _IP), ("sigmag", _DP), ("sigmae", _DP), ("pi", _DP),
                ("mu", _DP), ("m0", _IP), ("eps_final", _DP),
                ("perm", _IP), ("u", _DP), ("z", _DP), ("mu_draw", _DP), ("sigg_unit", _DP),
                ("pi_unit", _DP), ("sige_unit", _DP), ("sigmag_init", _DP), ("num_first", _DP),
                ("n_log_checked", C.c_int64), ("max_log_relerr", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_dot.restype = C.c_double
        _lib.oracle_last_error.restype = C.c_char_p
    return _lib


def _dp(a):
    return a.ctypes.data_as(_DP)


def _ip(a):
    return a.ctypes.data_as(_IP)


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def decode_tables():
    a = np.empty(1024); b = np.empty(1024); na = np.empty(64)
    lib().oracle_decode_tables(_dp(a), _dp(b), _dp(na))
    return a, b, na


def read_phen(path: str, N: int):
    im4 = (N + 3) // 4
    eps = np.empty(im4 * 4); mask4 = np.empty(im4, dtype=np.uint8)
    nonas = C.c_int(); nas = C.c_int()
    rc = lib().oracle_read_phen(path.encode(), N, _dp(eps), _u8(mask4), C.byref(nonas), C.byref(nas))
    if rc != 0:
        raise RuntimeError(f"oracle_read_phen({path}) failed")
    return eps, mask4, nonas.value, nas.value


def marker_stats(bed: np.ndarray, N: int, mask4: np.ndarray, nonas: int):
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    M = bed.shape[0]
    mave = np.empty(M); msig = np.empty(M)
    lib().oracle_marker_stats(_u8(bed), N, M, _u8(np.ascontiguousarray(mask4)), nonas, _dp(mave), _dp(msig))
    return mave, msig


def dot(bedcol: np.ndarray, eps: np.ndarray, mave: float, msig: float) -> float:
    bedcol = np.ascontiguousarray(bedcol, dtype=np.uint8)
    return lib().oracle_dot(_u8(bedcol), bedcol.size, _dp(np.ascontiguousarray(eps)), C.c_double(mave), C.c_double(msig))


def update_eps(eps: np.ndarray, mask4: np.ndarray, bedcol: np.ndarray, dbeta: float, mave: float, msig: float):
    d3 = np.array([dbeta, mave, msig])
    lib().oracle_update_eps(_dp(eps), _u8(np.ascontiguousarray(mask4)), mask4.size,
                            _u8(np.ascontiguousarray(bedcol, dtype=np.uint8)), _dp(d3))
    return eps


def block_of_markers(Mt: int, nranks: int, rank: int):
    S = C.c_int(); M = C.c_int(); Mm = C.c_int()
    lib().oracle_block_of_markers(Mt, nranks, rank, C.byref(S), C.byref(M), C.byref(Mm))
    return S.value, M.value, Mm.value


def gibbs(bed, eps0, mask4, nonas, group_index, cva, *, N, R=1, nrep=None, iterations=1, shuffle=True,
          rng_mode=1, seed=0, replay_dir=None, sync_rate=1):
    """Run the restated Bayes::process.  bed (Mt, mbytes) u8; eps0 (T, 4*im4); mask4 (T, im4);
    nonas (T,); group_index (Mt,); cva (G, K).  Returns a dict of history + dense variates."""
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    eps0 = np.ascontiguousarray(eps0, dtype=np.float64)
    mask4 = np.ascontiguousarray(mask4, dtype=np.uint8)
    nonas = np.ascontiguousarray(nonas, dtype=np.int32)
    group_index = np.ascontiguousarray(group_index, dtype=np.int32)
    cva = np.ascontiguousarray(cva, dtype=np.float64)
    Mt = bed.shape[0]; T = eps0.shape[0]; G, K = cva.shape
    if nrep is None:
        nrep = R if rng_mode == 0 else 1
    Mm = (Mt + R - 1) // R
    cfg = OracleCfg(N, Mt, T, G, K, R, nrep, iterations, int(shuffle), rng_mode, sync_rate, seed,
                    replay_dir.encode() if replay_dir else None)
    I = iterations
    im4 = (N + 3) // 4
    res = {
        "betas": np.zeros((I, T, Mt)), "comp": np.zeros((I, T, Mt), dtype=np.int32),
        "sigmag": np.zeros((I, T, G)), "sigmae": np.zeros((I, T)), "pi": np.zeros((I, T, G, K)),
        "mu": np.zeros((I, T)), "m0": np.zeros((I, T, G), dtype=np.int32), "eps_final": np.zeros((T, im4 * 4)),
        "perm": np.zeros((I, R, Mm), dtype=np.int32), "u": np.zeros((I, Mm, R, T)), "z": np.zeros((I, Mm, R, T)),
        "mu_draw": np.zeros((I, T)), "sigg_unit": np.zeros((I, T, G)), "pi_unit": np.zeros((I, T, G, K)),
        "sige_unit": np.zeros((I, T)), "sigmag_init": np.zeros((T, G)), "num_first": np.zeros((Mm, R, T)),
    }
    out = OracleOut()
    for name, _t in OracleOut._fields_:
        if name in res:
            a = res[name]
            setattr(out, name, _ip(a) if a.dtype == np.int32 else _dp(a))
    rc = lib().oracle_gibbs(C.byref(cfg), _u8(bed), _dp(eps0), _u8(mask4), _ip(nonas), _ip(group_index), _dp(cva),
                            C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle_gibbs rc={rc}: {lib().oracle_last_error().decode()}")
    res["n_log_checked"] = out.n_log_checked
    res["max_log_relerr"] = out.max_log_relerr
    return res


def predict(bed, mask4, nonas, y, mave, msig, beta_hist, *, N, R=1, keep=None):
    """Bayes::predict (bayes.cpp:14-284) for one trait under R ranks.  Returns g, beta, tdist, se, pval, sigma."""
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    Mt = bed.shape[0]
    im4 = (N + 3) // 4
    beta_hist = np.ascontiguousarray(beta_hist, dtype=np.float64).reshape(-1, Mt)
    y = np.ascontiguousarray(y, dtype=np.float64)
    mave = np.ascontiguousarray(mave, dtype=np.float64); msig = np.ascontiguousarray(msig, dtype=np.float64)
    mask4 = np.ascontiguousarray(mask4, dtype=np.uint8)
    k8 = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    g = np.empty(4 * im4); beta = np.empty(Mt); tdist = np.empty(Mt); se = np.empty(Mt); pval = np.empty(Mt)
    sigma = np.empty(R)
    f = lib().oracle_predict
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 7
    rc = f(N, Mt, R, bed.ctypes.data, mask4.ctypes.data, int(nonas), y.ctypes.data, mave.ctypes.data, msig.ctypes.data,
           beta_hist.ctypes.data, beta_hist.shape[0], None if k8 is None else k8.ctypes.data,
           g.ctypes.data, beta.ctypes.data, tdist.ctypes.data, se.ctypes.data, pval.ctypes.data, sigma.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_predict failed ({rc}): {lib().oracle_last_error().decode()}")
    return {"g": g, "beta": beta, "tdist": tdist, "se": se, "pval": pval, "sigma": sigma}


# ---------------------------------------------------------------- reference binary + its files

def have_reference() -> bool:
    return os.path.exists(REF_BIN)


def run_reference(workdir, bed, dim, phens, gri, grm, out_dir, *, iterations, seed=171014, nranks=1,
                  shuffle=1, threads=1, log_dir=None, extra=(), timeout=600):
    """Run oracle/_ref/gmrm_ref (the unmodified reference) with the reference's own flags."""
    # MALLOC_PERTURB_=255 makes glibc hand out zero-filled blocks: the reference leaves the pad slots of
    # epsilon_ uninitialised when N % 4 != 0 (phenotype.cpp:30 vs 653) and then multiplies them into every
    # dot product (bayes.cpp:756-764); zero is what a fresh process sees and what the oracle assumes.
    env = dict(os.environ, GMRM_SHIM_NRANKS=str(nranks), OMP_NUM_THREADS=str(threads), MALLOC_PERTURB_="255")
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
        env["GMRM_RNG_LOG_DIR"] = log_dir
    cmd = [REF_BIN, "--bed-file", bed, "--dim-file", dim, "--phen-files", ",".join(phens),
           "--group-index-file", gri, "--group-mixture-file", grm, "--shuffle-markers", str(shuffle),
           "--seed", str(seed), "--iterations", str(iterations), "--out-dir", out_dir, *extra]
    p = subprocess.run(cmd, cwd=workdir, env=env, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference failed ({p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    return p.stdout


def read_bet(path):
    """xfiles.hpp:24-37: uint32 Mt, then per saved iteration uint32 it + Mt doubles."""
    raw = np.fromfile(path, dtype=np.uint8)
    Mt = int(raw[:4].view(np.uint32)[0])
    rec = 4 + 8 * Mt
    n = (raw.size - 4) // rec
    its, vals = [], []
    for i in range(n):
        o = 4 + i * rec
        its.append(int(raw[o:o + 4].view(np.uint32)[0]))
        vals.append(raw[o + 4:o + rec].view(np.float64).copy())
    return np.array(its), np.stack(vals) if vals else np.zeros((0, Mt))


def read_cpn(path):
    raw = np.fromfile(path, dtype=np.uint8)
    Mt = int(raw[:4].view(np.uint32)[0])
    rec = 4 + 4 * Mt
    n = (raw.size - 4) // rec
    its, vals = [], []
    for i in range(n):
        o = 4 + i * rec
        its.append(int(raw[o:o + 4].view(np.uint32)[0]))
        vals.append(raw[o + 4:o + rec].view(np.int32).copy())
    return np.array(its), np.stack(vals) if vals else np.zeros((0, Mt), dtype=np.int32)


def read_mlma(path):
    """bayes.cpp:230-236: "%20s %8d %8d %20.15f %20.15f %20.15f %20.15f\n" = 123 bytes per kept marker:
    id, index in the .bim, index in the reference .bim, beta, t, se, p."""
    rows = []
    with open(path) as f:
        for line in f:
            v = line.split()
            if len(v) == 7:
                rows.append((v[0], int(v[1]), int(v[2]), float(v[3]), float(v[4]), float(v[5]), float(v[6])))
    return rows


def read_csv(path):
    """xfiles.cpp:17-42: it, G, sigmaG[G], sigmaE, h2, m0_sum, G, K, pi[G*K]."""
    rows = []
    with open(path) as f:
        for line in f:
            v = [x.strip() for x in line.strip().split(",")]
            if len(v) < 3:
                continue
            it, G = int(v[0]), int(v[1])
            sg = np.array([float(x) for x in v[2:2 + G]])
            sige, h2, m0s, G2, K = float(v[2 + G]), float(v[3 + G]), int(v[4 + G]), int(v[5 + G]), int(v[6 + G])
            pi = np.array([float(x) for x in v[7 + G:7 + G + G * K]]).reshape(G, K)
            rows.append({"it": it, "sigmag": sg, "sigmae": sige, "h2": h2, "m0_sum": m0s, "pi": pi})
    return rows


def load_inputs(bed_path, dim_path, phen_paths, gri_path, grm_path):
    """Inputs as the reference reads them, for feeding oracle.gibbs()."""
    N, Mt = (int(x) for x in open(dim_path).read().split()[:2])
    mbytes = (N + 3) // 4
    raw = np.fromfile(bed_path, dtype=np.uint8)
    bed = raw[3:3 + Mt * mbytes].reshape(Mt, mbytes)           # 3 magic bytes skipped, bayes.cpp:882
    eps, masks, nonas = [], [], []
    for p in phen_paths:
        e, m, n, _ = read_phen(p, N)
        eps.append(e); masks.append(m); nonas.append(n)
    gi = np.array([int(l.split()[1]) for l in open(gri_path) if l.strip()], dtype=np.int32)
    cva = np.array([[float(x) for x in l.split()] for l in open(grm_path) if l.strip()])
    return {"N": N, "Mt": Mt, "bed": bed, "eps0": np.stack(eps), "mask4": np.stack(masks),
            "nonas": np.array(nonas, dtype=np.int32), "group_index": gi, "cva": cva}


def write_replay_file(path, res, iterations=None):
    """The variates of an oracle.gibbs(..., rng_mode=0, replay_dir=...) run in the layout gmrm_b200_cli --replay-file reads
    (gmrm_b200/csrc/host/host.hpp, struct Replay): header, sigmag_init, then the arrays of gmrm_replay per iteration."""
    I, R, Mm = res["perm"].shape
    T = res["u"].shape[3]
    G, K = res["pi_unit"].shape[2], res["pi_unit"].shape[3]
    n = I if iterations is None else iterations
    with open(path, "wb") as f:
        f.write(b"GMRMRPL1")
        f.write(np.array([R, Mm, T, G, K, n], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(res["sigmag_init"], dtype=np.float64).tobytes())
        for i in range(n):
            f.write(np.ascontiguousarray(res["perm"][i], dtype=np.int32).tobytes())
            for name in ("u", "z", "mu_draw", "sigg_unit", "pi_unit", "sige_unit"):
                f.write(np.ascontiguousarray(res[name][i], dtype=np.float64).tobytes())
