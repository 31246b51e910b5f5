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
