"""One-time ingestion through the executable, from a real PLINK .bed file (the reference's
Bayes::load_genotype, src/bayes.cpp:867-900): writes a UKB-shaped data set (N = 458,000, M markers; no missing genotypes,
one block of random columns repeated -- only the byte rate matters here) to a scratch directory, runs
`gmrm_b200_cli` on it for two iterations and reports the executable's own "time to load genotype data" line as GB/s.
Usage: python tools/cli_ingest_probe.py [--markers 400000] [--dir /tmp/ingest]"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--markers", type=int, default=400000)
    ap.add_argument("--individuals", type=int, default=458000)
    ap.add_argument("--dir", default="/tmp/gmrm_ingest")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    N, M = a.individuals, a.markers
    mbytes = (N + 3) // 4
    os.makedirs(a.dir, exist_ok=True)
    stem = os.path.join(a.dir, "ukb")
    rng = np.random.default_rng(1)
    # bytes without the missing code 01: every 2-bit field from {00, 10, 11}
    fields = np.array([0, 2, 3], dtype=np.uint8)
    tab = np.array([fields[e % 3] | (fields[(e // 3) % 3] << 2) | (fields[(e // 9) % 3] << 4) | (fields[e // 27] << 6) for e in range(81)], dtype=np.uint8)
    block_markers = min(M, 4096)
    block = tab[rng.integers(0, 81, size=(block_markers, mbytes), dtype=np.uint8)]
    if N % 4:
        block[:, -1] &= (1 << (2 * (N % 4))) - 1
    t0 = time.time()
    with open(stem + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        done = 0
        while done < M:
            n = min(block_markers, M - done)
            f.write(block[:n].tobytes())
            done += n
    with open(stem + ".dim", "w") as f:
        f.write(f"{N} {M}\n")
    with open(stem + ".gri", "w") as f:
        f.write("".join(f"{j} 0\n" for j in range(M)))
    with open(stem + ".grm", "w") as f:
        f.write("0.00000 0.00010 0.00100 0.01000\n")
    y = rng.normal(size=N)
    with open(stem + ".phen", "w") as f:
        f.write("".join(f"{i + 1} {i + 1} {float(y[i])!r}\n" for i in range(N)))
    t_write = time.time() - t0
    size = os.path.getsize(stem + ".bed")
    cmd = [os.path.join(ROOT, "gmrm_b200", "gmrm_b200_cli"), "--bed-file", stem + ".bed", "--dim-file", stem + ".dim", "--phen-files", stem + ".phen",
           "--group-index-file", stem + ".gri", "--group-mixture-file", stem + ".grm", "--iterations", "2", "--seed", "1", "--out-dir",
           os.path.join(a.dir, "out"), "--gpus", str(a.gpus)]
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.time() - t0
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith(("INFO", "RESULT", "WARNING", "FATAL"))]
    load = next((float(m.group(1)) for ln in lines if (m := re.search(r"time to load genotype data = ([0-9.]+)", ln))), None)
    out = {"bed_bytes": size, "N": N, "M": M, "gpus": a.gpus, "write_files_s": round(t_write, 1), "cli_rc": p.returncode, "cli_wall_s": round(wall, 2),
           "load_genotype_s": load, "load_gbs": None if not load else round(size / load / 1e9, 2),
           "full_matrix_s_at_this_rate": None if not load else round(114.5e9 / (size / load), 1),
           "what": "gmrm_b200_cli reading a real .bed file (page cache warm from the write) into pinned host memory and uploading it: "
                   "file read + H2D + transcode + missing lists, the executable's own timer",
           "cli_lines": lines[:14]}
    print(json.dumps(out))
    for fn in (".bed", ".gri", ".phen"):
        try:
            os.remove(stem + fn)
        except OSError:
            pass
    return 0 if p.returncode == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
