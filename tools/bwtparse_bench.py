#!/usr/bin/env python3
"""Measurement of the bwtparse stage on BASELINE config 2: the 4 GB pan-genome is
parsed on the GPU, pfpb200_bwtparse_device runs on the parse where it lies in HBM (CUDA-event
times), and the UNMODIFIED bwtparse of the reference is timed on the host cores on the parse of
the first --cpu-haplotypes haplotypes (a bounded sample; its files come from the same GPU parse).
Prints one JSON line.  usage: bwtparse_bench.py [--haplotypes 100] [--base-len 40000000] [--steps 5]"""
import argparse, json, os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--haplotypes", type=int, default=100)
ap.add_argument("--base-len", type=int, default=40_000_000)
ap.add_argument("--cpu-haplotypes", type=int, default=10)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
pkg = g.load_package()
sc = pkg.pfp.Scanner(0)
text = pkg.synth.pangenome_text(a.base_len, a.haplotypes, 2, device="cuda")
torch.cuda.synchronize()
out = sc.parse_device(text, 10, 100, sai=True)
n = out.n_phrases
res = []
for _ in range(a.steps + 1):
    r = sc.bwtparse_device(out.parse, n, out.last, out.sai)
    res.append(r.as_dict())
res = res[1:]
med = sorted(res, key=lambda d: d["ms_total"])[len(res) // 2]
line = {"stage": "bwtparse (SA of the parse + BWT + ilist/bwlast/bwsai)", "n_phrases": n, "alphabet": med["alphabet"],
        "rounds": med["rounds"], "launches": med["launches"], "ms_total": med["ms_total"], "ms_sa": med["ms_sa"],
        "ms_lists": med["ms_lists"], "phrases_per_s": n / med["ms_total"] * 1e3,
        "alg_bytes": 4 * n + n + 5 * n + 4 * (n + 1) + (n + 1) + 5 * (n + 1),
        "workload": f"{a.haplotypes} haplotypes x {a.base_len} bp, w=10 p=100 -s"}
line["alg_GBps"] = line["alg_bytes"] / med["ms_total"] / 1e6
from oracle import bwtparse_oracle as bo
if bo.have_reference() and a.cpu_haplotypes:
    sub = text[: text.numel() * a.cpu_haplotypes // a.haplotypes]
    o2 = sc.parse_device(sub, 10, 100, sai=True)
    f = sc.fetch(o2)
    d = tempfile.mkdtemp(prefix="bpbench_")
    base = os.path.join(d, "x")
    for ext in ("parse", "last", "sai", "occ"):
        open(base + "." + ext, "wb").write(getattr(f, ext))
    t0 = time.time()
    subprocess.run([bo.REF_BWTPARSE, base, "-s"], check=True, stdout=subprocess.DEVNULL)
    sec = time.time() - t0
    r2 = sc.bwtparse_device(o2.parse, o2.n_phrases, o2.last, o2.sai)
    same = all(open(base + "." + e, "rb").read() == sc.to_host(p, k * r2.n_out)
               for e, p, k in (("ilist", r2.ilist, 4), ("bwlast", r2.bwlast, 1), ("bwsai", r2.bwsai, 5)))
    shutil.rmtree(d, ignore_errors=True)
    line["cpu_baseline"] = {"kind": "reference", "cores": 1, "sample": f"bwtparse -s (unmodified, oracle/_ref) on the parse of the first {a.cpu_haplotypes} haplotypes: {o2.n_phrases} phrases, wall clock {sec:.2f} s",
                            "phrases_per_s": o2.n_phrases / sec, "gpu_ms_same_sample": r2.ms_total, "outputs_identical": same}
print(json.dumps(line))
