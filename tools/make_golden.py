#!/usr/bin/env python3
"""Generate tests/golden/golden.npz from the UNMODIFIED reference scanner.

Runs oracle/_ref/newscanNT.x (compiled from /root/reference by oracle/Makefile) on a fixed
list of small seeded inputs and stores input + the five output files of every case.  The
fixtures pin oracle/pfp_oracle.c (tests/test_oracle_golden.py) and, through it, the CUDA path.

Usage (build container only; needs oracle/_ref):  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pfp_oracle as orc  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

synth = load_package().synth


def rnd_dna(n, seed):
    return synth.random_dna(n, seed).numpy().tobytes()


def cases():
    out = []

    def add(name, data, w=10, p=100, fasta=False):
        out.append(dict(name=name, data=data, w=w, p=p, fasta=fasta))

    # (w,p) sweep on random DNA (SURVEY appendix A list)
    for (w, p) in [(10, 100), (6, 50), (16, 500), (32, 1000), (4, 10), (10, 50), (16, 100)]:
        add(f"dna20k_w{w}_p{p}", rnd_dna(20000, 11), w, p)
    # first-window trigger: hash(ACGACGCGCT) % 100 == 0  
    add("first_window", b"ACGACGCGCT" + rnd_dna(5000, 12))
    # degenerate lengths
    add("n0", b"")
    add("n3_lt_w", b"ACG")
    add("n10_eq_w", b"ACGTTGCAAC")
    add("n10_eq_w_trigger", b"ACGACGCGCT")
    add("n11", b"ACGTTGCAACG")
    # single giant phrase / trigger-everywhere
    add("all_A", b"A" * 3000)
    add("all_N", b"N" * 3000)
    add("all_A_w4_p10", b"A" * 500, 4, 10)
    for c in range(65, 91):   # find a letter whose constant window triggers at w=4,p=10
        if orc.window_hash(bytes([c]) * 4) % 10 == 0:
            add("const_trigger_everywhere", bytes([c]) * 400, 4, 10)
            break
    # all byte values 3..255
    rng = np.random.default_rng(13)
    add("bytes_3_255", rng.integers(3, 256, 30000, dtype=np.uint8).tobytes(), 10, 100)
    add("bytes_3_255_w5_p17", rng.integers(3, 256, 30000, dtype=np.uint8).tobytes(), 5, 17)
    # text ending exactly on a trigger
    t = rnd_dna(8000, 14)
    e = orc.triggers(t, 10, 100)
    add("ends_on_trigger", t[: int(e[-1]) + 1])
    add("ends_one_past_trigger", t[: int(e[-1]) + 2])
    # invalid byte truncates the input (newscan.cpp:364)
    add("invalid_byte", t[:3000] + b"\x01" + t[3000:4000])
    add("invalid_byte_at_0", b"\x02ACGT")
    # repetitive pan-genome mini, plain and as FASTA
    recs = [r.numpy() for r in synth.pangenome_records(3000, 12, 5)]
    add("pangenome_plain", b"".join(r.tobytes() for r in recs))
    add("pangenome_fasta", synth.to_fasta(recs), fasta=True)
    add("pangenome_fasta_w6_p50", synth.to_fasta(recs), 6, 50, fasta=True)
    # FASTA corner cases (SURVEY appendix B4 + kseq.h:177-218)
    s1, s2 = rnd_dna(700, 15), rnd_dna(333, 16)
    add("fasta_crlf", synth.to_fasta([np.frombuffer(s1, np.uint8)], newline=b"\r\n"), fasta=True)
    add("fasta_lower_N", b">a desc\n" + s1[:300].lower() + b"\nNNNNnnnnNN\n" + s1[300:] + b"\n",
        fasta=True)
    add("fasta_empty_record", b">e1\n>e2\n" + s1 + b"\n>e3\n\n\n" + s2 + b"\n>e4\n", fasta=True)
    add("fasta_no_trailing_nl", b">x\n" + s1 + b"\n>y\n" + s2, fasta=True)
    add("fasta_blank_lines", b"\n\n>x\n\n" + s1[:100] + b"\n\n\n" + s1[100:] + b"\n\n", fasta=True)
    add("fasta_junk_before_header", b"junk line\nACGT\n>x c\n" + s1 + b"\n", fasta=True)
    add("fasta_header_only", b">only", fasta=True)
    add("fasta_empty_file", b"", fasta=True)
    add("fasta_gt_inside_line", b">x\nACGT>ACGT\n" + s1 + b"\n", fasta=True)
    add("fasta_high_byte", b">x\n" + s1[:200] + b"\xc3\xa9" + s1[200:] + b"\n", fasta=True)
    add("fasta_ff_byte", b">x\n" + s1[:200] + b"\xff" + s1[200:] + b"\n", fasta=True)
    add("fasta_ctrl_byte", b">x\n" + s1[:150] + b"\x02" + s1[150:] + b"\n>y\n" + s2 + b"\n", fasta=True)
    add("fastq_two_reads", b"@r1\n" + s1[:120] + b"\n+\n" + b"I" * 120 + b"\n@r2 c\n" +
        s2[:90] + b"\n+r2\n" + b"#" * 90 + b"\n", fasta=True)
    add("fastq_qual_with_at", b"@r1\n" + s1[:50] + b"\n+\n" + b"@" * 50 + b"\n@r2\n" + s2[:60] +
        b"\n+\n" + b">" * 60 + b"\n", fasta=True)
    return out


def main():
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle ref"
    arrays, meta = {}, []
    for c in cases():
        ref = orc.run_reference(c["data"], c["w"], c["p"], fasta=c["fasta"])
        name = c["name"]
        arrays[name + "/input"] = np.frombuffer(c["data"], np.uint8)
        for ext in ("dict", "occ", "parse", "last", "sai"):
            arrays[f"{name}/{ext}"] = np.frombuffer(getattr(ref, ext), np.uint8)
        meta.append(dict(name=name, w=c["w"], p=c["p"], fasta=c["fasta"],
                         n_phrases=ref.n_phrases, n_distinct=ref.n_distinct))
        print(f"{name:32s} n={len(c['data']):6d} phrases={ref.n_phrases:5d} distinct={ref.n_distinct:5d}")
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    np.savez_compressed(os.path.join(gdir, "golden.npz"), **arrays)
    with open(os.path.join(gdir, "cases.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py",
                       reference="alshai/Big-BWT newscanNT.x -s [-f] (oracle/_ref, unmodified)",
                       cases=meta), f, indent=1)
    print("wrote", len(meta), "cases")


if __name__ == "__main__":
    main()
