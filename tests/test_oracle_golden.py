"""Pins oracle/pfp_oracle.c against the reference's own outputs.

The golden fixtures were produced by the unmodified newscanNT.x (tools/make_golden.py); when
oracle/_ref is present (build container, GPU box) the reference is also run live."""
import os

import numpy as np
import pytest

from conftest import assert_same_files, golden, golden_dicz, golden_names
from oracle import pfp_oracle as orc


def _text_of(case):
    if case["fasta"]:
        text, _ = orc.fasta_extract(case["input"])
        return text
    return case["input"]


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    c = golden().case(name)
    got = orc.parse(_text_of(c), c["w"], c["p"])
    assert_same_files(got, c, name)
    assert got.n_phrases == c["n_phrases"] and got.n_distinct == c["n_distinct"]


@pytest.mark.parametrize("name", sorted(golden_dicz()))
def test_oracle_dicz_matches_golden(name):
    """-c: the oracle's .dicz rule against the file newscanNT.x -c wrote (newscan.cpp:410-413)."""
    meta, dicz = golden_dicz()[name]
    c = golden().case(name)
    assert orc.dicz_of(orc.parse(_text_of(c), c["w"], c["p"]).dict, meta["w"]) == dicz


def test_kr_hash_kat():
    # first-window example of SURVEY appendix B2: window hash 48505500 for w=10
    assert orc.window_hash(b"ACGACGCGCT") == 48505500
    # phrase hash definition (newscan.cpp:229-239): base-256 integer mod the 55-bit prime
    s = b"\x02ACGACGCGCT"
    assert orc.kr_hash(s) == int.from_bytes(s, "big") % 27162335252586509


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,w,p", [(101, 10, 100), (102, 4, 10), (103, 12, 37), (104, 31, 64),
                                       (105, 8, 1000)])
def test_oracle_matches_live_reference(pkg, seed, w, p):
    text = pkg.synth.pangenome_text(5000, 8, seed).numpy().tobytes()
    ref = orc.run_reference(text, w, p)
    assert_same_files(orc.parse(text, w, p), ref, f"seed{seed}")


@pytest.mark.skipif(not orc.have_ref("pscan.x"), reason="oracle/_ref not built")
def test_threaded_rule_matches_pscan():
    # helper-thread scanners skip a trigger at the very first window
    rng = np.random.default_rng(7)
    text = b"ACGACGCGCT" + rng.choice(np.frombuffer(b"ACGT", np.uint8), 20000).tobytes()
    ref = orc.run_reference(text, 10, 100, exe="pscan.x", threads=2)
    got = orc.parse(text, 10, 100, flags=orc.THREADED_RULE)
    assert_same_files(got, ref, "pscan -t 2")
    seq = orc.parse(text, 10, 100)
    assert seq.n_phrases == got.n_phrases + 1


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
def test_fasta_extract_matches_live_reference(pkg):
    recs = [r.numpy() for r in pkg.synth.pangenome_records(2000, 5, 9)]
    fa = pkg.synth.to_fasta(recs, width=70)
    text, trunc = orc.fasta_extract(fa)
    assert not trunc and text == b"".join(r.tobytes() for r in recs)
    assert_same_files(orc.parse(text), orc.run_reference(fa, fasta=True), "fasta live")


# ---- the bwtparse stage: oracle/bwtparse_oracle.py against the reference binary ----
def _bwtparse_golden():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_bwtparse.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {f: (z[f"{n}/{f}"].tobytes() if z[f"{n}/{f}"].ndim else int(z[f"{n}/{f}"]))
                for f in ("w", "p", "parse", "last", "sai", "occ", "ilist", "bwlast", "bwsai")} for n in names}


def test_bwtparse_oracle_matches_golden():
    from oracle import bwtparse_oracle as bo
    cases = _bwtparse_golden()
    assert len(cases) >= 6
    for name, c in cases.items():
        il, bl, bs = bo.bwtparse(c["parse"], c["last"], c["sai"])
        assert il == c["ilist"] and bl == c["bwlast"] and bs == c["bwsai"], name
        il2, bl2, bs2 = bo.bwtparse(c["parse"], c["last"], None)
        assert il2 == c["ilist"] and bl2 == c["bwlast"] and bs2 == b"", name


@pytest.mark.skipif(not orc.have_ref("bwtparse"), reason="oracle/_ref not built")
def test_bwtparse_oracle_matches_live_reference(pkg):
    from oracle import bwtparse_oracle as bo
    text = pkg.synth.pangenome_text(30000, 10, 77).numpy().tobytes()
    f = orc.parse(text, 8, 40)
    for nseg in (0, 3):
        want = bo.run_reference(f.parse, f.last, f.sai, f.occ, nseg=nseg)
        assert bo.bwtparse(f.parse, f.last, f.sai) == want, f"nseg={nseg}"


# ---- the pfbwt stage: oracle/pfbwt_oracle.py against the reference binary ----
def _pfbwt_golden():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_pfbwt.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {f: (z[f"{n}/{f}"].tobytes() if z[f"{n}/{f}"].ndim else int(z[f"{n}/{f}"]))
                for f in ("w", "p", "text", "dict", "occ", "ilist", "bwlast", "bwsai", "bwt", "sa", "ssa", "esa")}
            for n in names}


def test_pfbwt_oracle_matches_golden():
    from oracle import pfbwt_oracle as po
    cases = _pfbwt_golden()
    assert len(cases) >= 6
    for name, c in cases.items():
        got = po.pfbwt(c["text"])
        for ext in ("bwt", "sa", "ssa", "esa"):
            assert got[ext] == c[ext], f"{name}: .{ext}"


def test_pfbwt_algorithm_restatement_matches_golden():
    """The walk over the sorted dictionary suffixes with the ilist merge (pfbwt.cpp:109-242), restated,
    gives the reference binary's .bwt and .sa from the reference's own intermediate files."""
    from oracle import pfbwt_oracle as po
    cases = _pfbwt_golden()
    for name in ("short_w4_p10", "low_complexity_w4_p11", "identical_copies_w10_p50", "random_30k_w4_p10"):
        c = cases[name]
        got = po.pfbwt_algorithm(c["dict"], c["occ"], c["ilist"], c["bwlast"], c["bwsai"], c["w"])
        assert got["bwt"] == c["bwt"], f"{name}: .bwt"
        assert got["sa"] == c["sa"], f"{name}: .sa"
