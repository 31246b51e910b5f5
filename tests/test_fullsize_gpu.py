"""BASELINE.json's full-size configurations on one GPU.  Byte-exact where the build container
could run the unmodified newscanNT.x on the same full-size text (tests/golden/fullsize_sha256.json,
tools/make_fullsize_digests.py): sha256 of each of the five streams == the reference's.  Plus
size-independent properties (tools/fullsize_check.py: counts, bincount(.parse) == .occ, .sai/.last
against the text, every phrase end a trigger and every trigger of a 64 MB slice a phrase end,
EVERY adjacent dictionary pair in order, unparse of a 200 000-phrase prefix)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _need(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free device memory")


@pytest.fixture(scope="module")
def sc(pkg):
    s = pkg.pfp.Scanner(0)
    yield s
    s.close()


def test_config2_pangenome_4gb_and_sweep_corners(pkg, sc):
    """config 2 (100 haplotypes x 40 Mbp, w=10 p=100) and the corners of config 5's sweep."""
    import fullsize_check as fc
    _need(40)
    text = pkg.synth.pangenome_text(40_000_000, 100, 2, device="cuda")
    for (w, p, key) in [(10, 100, "config2 w10 p100"), (6, 50, "sweep w6 p50"), (16, 500, "sweep w16 p500"),
                        (32, 1000, "sweep w32 p1000"), (6, 1000, None), (32, 50, None)]:
        assert fc.check_case(sc, text, w, p, f"pangenome 4 GB w{w} p{p}", digest_key=key,
                             text_key="pangenome" if key else None,
                             pipeline=(w, p) == (10, 100)), f"w={w} p={p}"     # config 2: also bwtparse + pfbwt -S


def test_config4_random_8gb(pkg, sc):
    """config 4 at full size: 8 GB of uniform random ACGT, nearly every phrase distinct."""
    import fullsize_check as fc
    _need(100)
    text = pkg.synth.random_dna(8_000_000_000, 4, device="cuda")
    assert fc.check_case(sc, text, 10, 100, "config4 random ACGT 8 GB", digest_key="config4 random 8 GB w10 p100",
                         text_key="random")


def test_dict_order_check_kernel(pkg, sc):
    """The all-pairs order check itself: 0 on a sorted dictionary, counts every violation."""
    words = sorted({pkg.synth.random_dna(int(n), 500 + i).numpy().tobytes()
                    for i, n in enumerate([12, 12, 40, 7, 300, 12, 1, 2, 64, 64, 65])} | {b"AC", b"ACG", b"ACGT"})

    def bad(ws):
        d = torch.from_numpy(__import__("numpy").frombuffer(b"".join(x + b"\x01" for x in ws) + b"\x00", dtype="uint8").copy()).cuda()
        seps = torch.nonzero(d == 1).flatten()
        torch.cuda.synchronize()                         # the library runs on its own stream
        return sc.check_dict_order(d, seps)
    assert bad(words) == 0
    assert bad(words[::-1]) == len(words) - 1
    assert bad(words[:3] + [words[2]] + words[3:]) == 1          # a duplicate is not strictly increasing
    sw = list(words)
    sw[4], sw[5] = sw[5], sw[4]
    assert bad(sw) >= 1
