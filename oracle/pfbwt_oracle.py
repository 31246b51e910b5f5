"""CPU checker for the last stage (reference pfbwt.cpp): the BWT and suffix array of the text.

TEST INFRASTRUCTURE ONLY: imported by tests/ as the checker of pfpb200_pfbwt_*; nothing under
big-bwt_b200/ may import this.  It restates the RESULT the reference defines rather than its
merge loop: pfbwt*.x writes the BWT of T$ ($ = 0, smallest; .bwt, n + 1 chars), with -S the suffix
array without the entry of the $ suffix (.sa, n values of 5 bytes, pfbwt.cpp:157-163), with -s / -e
the (position, value) pairs at the starts / ends of the BWT's runs (.ssa / .esa, :165-193), where
the value at position 0 is the text length (:182).  Pinned: tests/golden/golden_pfbwt.npz holds
the outputs of the UNMODIFIED chain newscanNT.x -> bwtparse -> pfbwtNT.x (tools/make_golden_pfbwt.py)
and tests/test_oracle_golden.py::test_pfbwt_oracle_* compares byte for byte.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile

import numpy as np

from oracle.bwtparse_oracle import suffix_array

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PFBWT = os.path.join(HERE, "_ref", "pfbwtNT.x")
SABYTES = 5                                  # utils.h:12


def _put5(vals: np.ndarray) -> bytes:
    v = vals.astype(np.uint64)
    return np.stack([(v >> np.uint64(8 * j)) & np.uint64(255) for j in range(SABYTES)], axis=1).astype(np.uint8).tobytes()


def pfbwt(text: bytes):
    """{'bwt', 'sa', 'ssa', 'esa'} bytes as pfbwtNT.x writes them for the text (any w, p)."""
    t = np.frombuffer(text, dtype=np.uint8).astype(np.int64)
    n = t.size
    s = np.concatenate([t, [0]])                              # T$
    sa = suffix_array(s)                                      # sa[0] = n: the $ suffix
    bwt = np.where(sa == 0, 0, s[sa - 1]).astype(np.uint8)    # the char in front of T is the EOF char 0 (:127)
    pos = np.arange(n + 1)
    start = np.concatenate([[True], bwt[1:] != bwt[:-1]])
    end = np.concatenate([bwt[1:] != bwt[:-1], [True]])
    pair = lambda m: _put5(np.stack([pos[m], sa[m]], axis=1).reshape(-1))    # noqa: E731
    return {"bwt": bwt.tobytes(), "sa": _put5(sa[1:]), "ssa": pair(start), "esa": pair(end)}


def run_reference(files: dict, w: int, flags=("-S",)):
    """The unmodified pfbwtNT.x on {.dict .occ .ilist .bwlast .bwsai} -> dict of its outputs."""
    d = tempfile.mkdtemp(prefix="pfbwt_ref_")
    try:
        base = os.path.join(d, "x")
        for ext, data in files.items():
            with open(base + "." + ext, "wb") as f:
                f.write(data)
        subprocess.run([REF_PFBWT, "-w", str(w), *flags, base], check=True, stdout=subprocess.DEVNULL)
        out = {}
        for ext in ("bwt", "sa", "ssa", "esa"):
            if os.path.exists(base + "." + ext):
                out[ext] = open(base + "." + ext, "rb").read()
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def have_reference() -> bool:
    return os.path.exists(REF_PFBWT)


def pfbwt_algorithm(dict_b: bytes, occ_b: bytes, ilist_b: bytes, bwlast_b: bytes, bwsai_b: bytes | None, w: int):
    """The reference's ALGORITHM (pfbwt.cpp bwt(), :109-242) restated for small inputs, next to the
    result definition above: sort the dictionary's suffixes (compute_dict_bwt_lcp, :483-515; here a
    plain sort of the byte strings up to their word's EndOfWord), walk them in order, skip those of
    length <= w (:152), emit for a whole word the .bwlast char of each of its occurrences (:154-203),
    for a group of equal proper suffixes of several words the chars in front of them merged by the
    words' positions in the BWT of the parse (fwrite_chars_same_suffix, :521-560 -- the heap merge
    is a sort of the (ilist position, char) pairs).  Returns {'bwt', 'sa'} (sa = b'' without bwsai)."""
    d = bytearray(dict_b)
    occ = np.frombuffer(occ_b, dtype=np.uint32).astype(np.int64)
    ilist = np.frombuffer(ilist_b, dtype=np.uint32).astype(np.int64)
    sai = None
    if bwsai_b:
        a = np.frombuffer(bwsai_b, dtype=np.uint8).reshape(-1, SABYTES).astype(np.uint64)
        sai = sum(a[:, j] << np.uint64(8 * j) for j in range(SABYTES)).astype(np.int64)
    assert d[0] == 2 and d[-1] == 0                          # :126, :496
    d[0] = 0                                                 # the EOF char of the final BWT (:127)
    istart = np.concatenate([[1], 1 + np.cumsum(occ)])       # :376-383: ilist[0] belongs to the end symbol
    assert istart[-1] == ilist.size
    # (suffix bytes, word id, start) of every dictionary suffix longer than w
    sufs, start, word = [], 0, 0
    for e in range(len(d)):
        if d[e] == 1:                                        # EndOfWord
            for t in range(start, e):
                if e - t > w:
                    sufs.append((bytes(d[t:e]) if t else bytes([2]) + bytes(d[1:e]), word, t, t == start))
            start, word = e + 1, word + 1
    assert word == occ.size                                  # :360
    sufs.sort(key=lambda s: (s[0], s[1]))                    # equal suffixes: the order inside a group is irrelevant
    bwt, sa = bytearray(), []
    i = 0
    while i < len(sufs):
        j = i
        while j < len(sufs) and sufs[j][0] == sufs[i][0]:
            j += 1
        group = sufs[i:j]
        length = len(group[0][0])
        pairs = []
        for _, wd, t, full in group:
            for k in range(istart[wd], istart[wd + 1]):
                pos = int(ilist[k])
                pairs.append((pos, bwlast_b[pos] if full else d[t - 1], wd, full))
        pairs.sort()                                         # the heap of ilist cursors (:536-558)
        for pos, ch, wd, full in pairs:
            bwt.append(ch)
            if sai is not None and not (full and wd == 0):   # no SA entry for the first word of the parse (:157-163)
                sa.append(int(sai[pos]) - length)
        i = j
    return {"bwt": bytes(bwt), "sa": _put5(np.array(sa, dtype=np.int64)) if sai is not None else b""}
