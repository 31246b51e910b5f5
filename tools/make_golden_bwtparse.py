#!/usr/bin/env python3
"""Golden vectors for the bwtparse stage: runs the UNMODIFIED reference chain in the build
container -- oracle/_ref/newscanNT.x -s on seeded texts, then oracle/_ref/bwtparse -s on its
files -- and stores inputs and outputs in tests/golden/golden_bwtparse.npz.
usage: python tools/make_golden_bwtparse.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pfp_oracle as orc
from oracle import bwtparse_oracle as bo
import __graft_entry__ as g

pkg = g.load_package()
cases = {}


def add(name, text, w, p):
    ref = orc.run_reference(text, w, p, sai=True)            # the five files of newscanNT.x -s
    il, bl, bs = bo.run_reference(ref.parse, ref.last, ref.sai, ref.occ)
    cases[name] = dict(w=w, p=p, parse=ref.parse, last=ref.last, sai=ref.sai, occ=ref.occ, ilist=il, bwlast=bl, bwsai=bs)
    print(name, "phrases", len(ref.parse) // 4, "ilist", len(il) // 4)


add("pangenome_20k_x6_w10_p100", pkg.synth.pangenome_text(20000, 6, seed=5).numpy().tobytes(), 10, 100)
add("pangenome_50k_x8_w6_p20", pkg.synth.pangenome_text(50000, 8, seed=6).numpy().tobytes(), 6, 20)
add("random_30k_w4_p10", pkg.synth.random_dna(30000, 7).numpy().tobytes(), 4, 10)
add("identical_copies_w10_p50", pkg.synth.random_dna(5000, 8).numpy().tobytes() * 12, 10, 50)
add("low_complexity_w4_p11", (b"ACACACGT" * 40 + pkg.synth.random_dna(300, 9).numpy().tobytes() + b"T" * 200) * 20, 4, 11)
add("short_w4_p10", pkg.synth.random_dna(400, 10).numpy().tobytes(), 4, 10)
flat = {}
for k, c in cases.items():
    for f, v in c.items():
        flat[f"{k}/{f}"] = np.frombuffer(v, dtype=np.uint8) if isinstance(v, (bytes, bytearray)) else np.int64(v)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_bwtparse.npz"), **flat)
print("written", len(cases), "cases")
