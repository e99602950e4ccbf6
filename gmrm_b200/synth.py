"""Synthetic PLINK-style inputs for tests and benches.

Genotypes follow example/data_sim.R:15 of the reference (Binomial(2, p) minor-allele counts),
phenotypes follow data_sim.R:17-41 (y = scale(X) b + e with a fixed number of causal markers
and heritability h2).  Files are written in exactly the formats the reference reads:

  .bed  3 magic bytes 6c 1b 01, then SNP-major, ceil(N/4) bytes per marker, individual 4i+k in
        bits 2k..2k+1 of byte i; codes 00 = dosage 2, 10 = dosage 1, 11 = dosage 0, 01 = missing
        (reference decode tables src/lut/mk_lut.cpp:25-32, loader src/bayes.cpp:867-900)
  .dim  "N M"                                     (src/dimensions.cpp:8-29)
  .phen "FID IID value", literal NA for missing   (src/phenotype.cpp:587-673)
  .gri  "label group" per marker, 0-based group   (src/bayes.cpp:830-853)
  .grm  G lines of K ascending mixture variances, the first 0.0 (src/options.cpp:222-286)
"""
from __future__ import annotations

import os
import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
# dosage (0,1,2) or 3 = missing -> 2-bit PLINK code
_CODE_OF = np.array([0b11, 0b10, 0b00, 0b01], dtype=np.uint8)


def pack_bed(dosage: np.ndarray) -> np.ndarray:
    """dosage: (M, N) uint8 with values 0,1,2 or 3 (missing) -> (M, ceil(N/4)) packed bytes.

    Pad slots of the last byte are 00, as PLINK writes them."""
    M, N = dosage.shape
    mbytes = (N + 3) // 4
    codes = np.zeros((M, mbytes * 4), dtype=np.uint8)
    codes[:, :N] = _CODE_OF[dosage]
    codes = codes.reshape(M, mbytes, 4)
    return (codes[:, :, 0] | (codes[:, :, 1] << 2) | (codes[:, :, 2] << 4) | (codes[:, :, 3] << 6)).astype(np.uint8)


def make_genotypes(N: int, M: int, seed: int = 1, maf_lo: float = 0.05, maf_hi: float = 0.5,
                   missing_rate: float = 0.0) -> np.ndarray:
    """(M, N) uint8 dosages, per-marker MAF ~ U(maf_lo, maf_hi), optional missing (value 3)."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(maf_lo, maf_hi, size=M)
    d = rng.binomial(2, p[:, None], size=(M, N)).astype(np.uint8)
    if missing_rate > 0:
        d[rng.random((M, N)) < missing_rate] = 3
    return d


def make_phenotypes(dosage: np.ndarray, n_traits: int, h2: float = 0.5, causal_frac: float = 0.25,
                    seed: int = 171014):
    """y = scale(X) b + e per trait; returns (y (T, N) float64, true_beta (T, M))."""
    M, N = dosage.shape
    rng = np.random.default_rng(seed)
    X = dosage.astype(np.float64)
    X[dosage == 3] = np.nan
    mu = np.nanmean(X, axis=1, keepdims=True)
    sd = np.nanstd(X, axis=1, ddof=1, keepdims=True)
    sd[sd == 0] = 1.0
    Z = np.where(np.isnan(X), 0.0, (X - mu) / sd)
    ncausal = max(1, int(round(causal_frac * M)))
    ys, betas = [], []
    for _ in range(n_traits):
        beta = np.zeros(M)
        idx = rng.choice(M, ncausal, replace=False)
        beta[idx] = rng.normal(0.0, np.sqrt(h2 / ncausal), size=ncausal)
        g = beta @ Z
        e = rng.normal(0.0, np.sqrt(max(1e-12, 1.0 - g.var())), size=N)
        ys.append(g + e)
        betas.append(beta)
    return np.stack(ys), np.stack(betas)


def write_dataset(outdir: str, N: int, M: int, n_traits: int = 1, n_groups: int = 1,
                  mixtures=(0.0, 1e-4, 1e-3, 1e-2), na_rate: float = 0.0, missing_rate: float = 0.0,
                  seed: int = 1, h2: float = 0.5, causal_frac: float = 0.25, stem: str = "syn"):
    """Write <stem>.bed/.dim/.gri/.grm and <stem>_t<k>.phen; returns a dict of paths + arrays."""
    os.makedirs(outdir, exist_ok=True)
    dosage = make_genotypes(N, M, seed=seed, missing_rate=missing_rate)
    bed = pack_bed(dosage)
    y, true_beta = make_phenotypes(dosage, n_traits, h2=h2, causal_frac=causal_frac, seed=171014 + seed)
    rng = np.random.default_rng(seed + 7)
    paths = {k: os.path.join(outdir, f"{stem}.{k}") for k in ("bed", "dim", "gri", "grm")}
    with open(paths["bed"], "wb") as f:
        f.write(BED_MAGIC)
        f.write(bed.tobytes())
    with open(paths["dim"], "w") as f:
        f.write(f"{N} {M}\n")
    groups = rng.integers(0, n_groups, size=M) if n_groups > 1 else np.zeros(M, dtype=np.int64)
    with open(paths["gri"], "w") as f:
        for j in range(M):
            f.write(f"{j} {int(groups[j])}\n")
    mixtures = np.asarray(mixtures, dtype=np.float64)
    with open(paths["grm"], "w") as f:
        for g in range(n_groups):
            # per-group scaling pattern of benchmarking/generate_group_files.pl:24
            row = mixtures * (1.0 + g)
            f.write(" ".join(f"{v:.5f}" if v < 1 else f"{v:.5f}" for v in row) + "\n")
    phen_paths, na_masks = [], []
    for t in range(n_traits):
        p = os.path.join(outdir, f"{stem}_t{t}.phen")
        na = rng.random(N) < na_rate if na_rate > 0 else np.zeros(N, dtype=bool)
        with open(p, "w") as f:
            for i in range(N):
                v = "NA" if na[i] else repr(float(y[t, i]))
                f.write(f"{i + 1} {i + 1} {v}\n")
        phen_paths.append(p)
        na_masks.append(na)
    paths["phen"] = phen_paths
    return {"paths": paths, "dosage": dosage, "bed": bed, "y": y, "na": np.stack(na_masks),
            "groups": groups, "true_beta": true_beta, "N": N, "M": M}
