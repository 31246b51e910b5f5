#!/usr/bin/env python3
"""Measurement of the whole pipeline in HBM on BASELINE config 2: text ->
parse -> bwtparse -> pfbwt, each stage's CUDA-event time, and the UNMODIFIED reference chain
(newscanNT.x files from the GPU parse -> bwtparse -> pfbwtNT.x) timed on the host cores on the first
--cpu-haplotypes haplotypes (a bounded sample), outputs compared byte for byte.
Prints one JSON line.  usage: pfbwt_bench.py [--haplotypes 100] [--base-len 40000000] [--sa]"""
import argparse, json, os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--haplotypes", type=int, default=100)
ap.add_argument("--base-len", type=int, default=40_000_000)
ap.add_argument("--cpu-haplotypes", type=int, default=3)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--sa", action="store_true", help="also the full suffix array (-S)")
a = ap.parse_args()
pkg = g.load_package()
P = pkg.pfp
sc = P.Scanner(0)
text = pkg.synth.pangenome_text(a.base_len, a.haplotypes, 2, device="cuda")
torch.cuda.synchronize()
flags = P.PFBWT_SA if a.sa else 0
runs = []
for _ in range(a.steps + 1):
    t0 = time.perf_counter()
    r, out, bp = sc.bwt_of_text(text, 10, 100, flags=flags)
    wall = time.perf_counter() - t0
    runs.append({"ms_parse": sc.stats.ms_total, "ms_bwtparse": bp.ms_total, "ms_pfbwt": r.ms_total, "ms_pfbwt_sort": r.ms_sa,
                 "ms_pfbwt_fill": r.ms_fill, "wall_ms": wall * 1e3, "rounds": r.rounds, "easy": r.easy, "hard": r.hard,
                 "launches": r.launches})
runs = runs[1:]
med = sorted(runs, key=lambda d: d["wall_ms"])[len(runs) // 2]
n = text.numel()
line = {"pipeline": "text in HBM -> parse -> bwtparse -> pfbwt -> BWT" + (" + SA" if a.sa else "") + " in HBM",
        "workload": f"{a.haplotypes} haplotypes x {a.base_len} bp, w=10 p=100", "n_text": n, "dict_bytes": out.dict_bytes,
        "n_phrases": out.n_phrases, **med, "gpu_ms_sum": med["ms_parse"] + med["ms_bwtparse"] + med["ms_pfbwt"],
        "text_GBps_wall": n / med["wall_ms"] / 1e6}
from oracle import bwtparse_oracle as bo, pfbwt_oracle as po
if po.have_reference() and a.cpu_haplotypes:
    sub = text[: n * a.cpu_haplotypes // a.haplotypes].clone()
    r2, o2, b2 = sc.bwt_of_text(sub, 10, 100, flags=flags)
    gpu_ms = sc.stats.ms_total + b2.ms_total + r2.ms_total
    bwt_gpu = sc.to_host(r2.bwt, r2.n_bwt)
    sa_gpu = sc.to_host(r2.sa, 5 * r2.n_sa) if a.sa else b""
    o2 = sc.parse_device(sub, 10, 100, sai=True)
    f = sc.fetch(o2)
    d = tempfile.mkdtemp(prefix="pfbench_")
    base = os.path.join(d, "x")
    for ext in ("dict", "parse", "last", "sai", "occ"):
        open(base + "." + ext, "wb").write(getattr(f, ext))
    t0 = time.time()
    subprocess.run([bo.REF_BWTPARSE, base, "-s"], check=True, stdout=subprocess.DEVNULL)
    t1 = time.time()
    subprocess.run([po.REF_PFBWT, "-w", "10", *(["-S"] if a.sa else []), base], check=True, stdout=subprocess.DEVNULL)
    t2 = time.time()
    same = open(base + ".bwt", "rb").read() == bwt_gpu and (not a.sa or open(base + ".sa", "rb").read() == sa_gpu)
    shutil.rmtree(d, ignore_errors=True)
    line["cpu_baseline"] = {"kind": "reference", "cores": 1,
                            "sample": f"bwtparse -s + pfbwtNT.x{' -S' if a.sa else ''} (unmodified, oracle/_ref) on the first {a.cpu_haplotypes} haplotypes ({sub.numel()} bytes of text)",
                            "bwtparse_s": round(t1 - t0, 2), "pfbwt_s": round(t2 - t1, 2),
                            "text_MBps_pfbwt": sub.numel() / (t2 - t1) / 1e6, "gpu_ms_same_sample_all_stages": gpu_ms,
                            "outputs_identical": same}
print(json.dumps(line))
