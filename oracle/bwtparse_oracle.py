"""CPU restatement of the reference's bwtparse.c (the stage after the parse) in numpy.

TEST INFRASTRUCTURE ONLY: imported by tests/ as the checker of pfpb200_bwtparse_*; nothing under
big-bwt_b200/ may import this.  Pinned: tests/golden/golden_bwtparse.npz holds the outputs of the
UNMODIFIED bwtparse binary (oracle/_ref/bwtparse, built from /root/reference/bwtparse.c +
gsa/gsacak.c) on the seeded cases of tools/make_golden_bwtparse.py, and
tests/test_oracle_golden.py::test_bwtparse_oracle_* compares this restatement with them byte for
byte (and with live runs of the binary when oracle/_ref is present).

Follows bwtparse.c:
  * T = parse symbols + the end symbol 0 (read_parse, :70-117; Text[n] = 0 at :114)
  * SA = suffix array of T[0..n] (compute_SA -> sacak_int, :162-174); here: a sort of the suffixes
    by prefix doubling on ranks (numpy lexsort), same order because 0 is unique and smallest
  * BWT / .bwlast / .bwsai (:243-272), inverted list by counting sort (:276-306)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BWTPARSE = os.path.join(HERE, "_ref", "bwtparse")
IBYTES = 5                                   # utils.h:10


def suffix_array(text: np.ndarray) -> np.ndarray:
    """Suffix array of an integer string whose last symbol is a unique smallest end symbol."""
    n = text.size
    rank = np.unique(text, return_inverse=True)[1].astype(np.int64)
    h = 1
    while True:
        nxt = np.zeros(n, dtype=np.int64)
        nxt[:n - h] = rank[h:] if h < n else 0
        order = np.lexsort((nxt, rank))
        key = rank[order] * (n + 1) + nxt[order]
        new = np.zeros(n, dtype=np.int64)
        new[order] = np.concatenate([[0], np.cumsum(key[1:] != key[:-1])])
        rank = new
        if rank.max() == n - 1:
            return np.argsort(rank).astype(np.int64)
        h *= 2


def bwtparse(parse: bytes, last: bytes, sai: bytes | None = None):
    """(.ilist, .bwlast, .bwsai) bytes of bwtparse.c for the given .parse / .last / .sai bytes."""
    t = np.frombuffer(parse, dtype=np.uint32).astype(np.int64)
    n = t.size
    assert n > 1                                              # bwtparse.c:241
    la = np.frombuffer(last, dtype=np.uint8)
    text = np.concatenate([t, [0]])
    sa = suffix_array(text)
    assert sa[0] == n and sa[1] == 0                          # :244, :250
    bwt = np.where(sa == 0, 0, text[sa - 1])                  # :246, :252, :268
    li = np.where(sa == 1, n - 1, sa - 2)                     # :259-262 (sa == 0: dummy, overwritten below)
    bwlast = np.where(sa == 0, 0, la[np.clip(li, 0, n - 1)]).astype(np.uint8)
    bwsai = b""
    if sai:
        sv = np.frombuffer(sai, dtype=np.uint8).reshape(n, IBYTES)
        out = sv[np.clip(sa - 1, 0, n - 1)].copy()
        out[sa == 0] = 0                                      # :256: dummy position of the end symbol
        bwsai = out.tobytes()
    ilist = np.argsort(bwt, kind="stable").astype(np.uint32)  # :294-298: positions by symbol, ascending
    return ilist.tobytes(), bwlast.tobytes(), bwsai


def run_reference(parse: bytes, last: bytes, sai: bytes | None, occ: bytes, nseg: int = 0):
    """The unmodified bwtparse binary on files holding the given streams -> (.ilist, .bwlast, .bwsai)."""
    d = tempfile.mkdtemp(prefix="bwtparse_ref_")
    try:
        base = os.path.join(d, "x")
        with open(base + ".parse", "wb") as f:
            f.write(parse)
        with open(base + ".occ", "wb") as f:
            f.write(occ)
        n = len(last)
        if nseg > 0:                                          # segment files <base>.<i>.last|sai (utils.c:44-56)
            cuts = [n * i // nseg for i in range(nseg + 1)]
            for i in range(nseg):
                with open(f"{base}.{i}.last", "wb") as f:
                    f.write(last[cuts[i]:cuts[i + 1]])
                if sai:
                    with open(f"{base}.{i}.sai", "wb") as f:
                        f.write(sai[IBYTES * cuts[i]:IBYTES * cuts[i + 1]])
        else:
            with open(base + ".last", "wb") as f:
                f.write(last)
            if sai:
                with open(base + ".sai", "wb") as f:
                    f.write(sai)
        cmd = [REF_BWTPARSE, base] + (["-s"] if sai else []) + (["-t", str(nseg)] if nseg else [])
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        rd = lambda ext: open(base + ext, "rb").read()        # noqa: E731
        return rd(".ilist"), rd(".bwlast"), rd(".bwsai") if sai else b""
    finally:
        shutil.rmtree(d, ignore_errors=True)


def have_reference() -> bool:
    return os.path.exists(REF_BWTPARSE)
