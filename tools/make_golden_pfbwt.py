#!/usr/bin/env python3
"""Golden vectors for the pfbwt stage: the UNMODIFIED reference chain in the build container --
oracle/_ref/newscanNT.x -s, bwtparse -s, pfbwtNT.x (-S, then -s -e) -- on seeded texts; inputs of
the last stage and its outputs go to tests/golden/golden_pfbwt.npz.
usage: python tools/make_golden_pfbwt.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pfp_oracle as orc
from oracle import bwtparse_oracle as bo
from oracle import pfbwt_oracle as po
import __graft_entry__ as g

pkg = g.load_package()
flat = {}


def add(name, text, w, p):
    ref = orc.run_reference(text, w, p, sai=True)
    il, bl, bs = bo.run_reference(ref.parse, ref.last, ref.sai, ref.occ)
    files = {"dict": ref.dict, "occ": ref.occ, "ilist": il, "bwlast": bl, "bwsai": bs}
    full = po.run_reference(files, w, ("-S",))
    samp = po.run_reference(files, w, ("-s", "-e"))
    assert full["bwt"] == samp["bwt"]
    c = dict(files, text=text, bwt=full["bwt"], sa=full["sa"], ssa=samp["ssa"], esa=samp["esa"])
    for f, v in c.items():
        flat[f"{name}/{f}"] = np.frombuffer(v, dtype=np.uint8)
    flat[f"{name}/w"] = np.int64(w)
    flat[f"{name}/p"] = np.int64(p)
    print(name, "n", len(text), "words", len(ref.occ) // 4, "bwt", len(full["bwt"]))


add("pangenome_20k_x6_w10_p100", pkg.synth.pangenome_text(20000, 6, seed=5).numpy().tobytes(), 10, 100)
add("pangenome_30k_x8_w6_p20", pkg.synth.pangenome_text(30000, 8, seed=6).numpy().tobytes(), 6, 20)
add("random_30k_w4_p10", pkg.synth.random_dna(30000, 7).numpy().tobytes(), 4, 10)
add("identical_copies_w10_p50", pkg.synth.random_dna(5000, 8).numpy().tobytes() * 12, 10, 50)
add("low_complexity_w4_p11", (b"ACACACGT" * 40 + pkg.synth.random_dna(300, 9).numpy().tobytes() + b"T" * 200) * 20, 4, 11)
add("short_w4_p10", pkg.synth.random_dna(400, 10).numpy().tobytes(), 4, 10)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_pfbwt.npz"), **flat)
print("written")
