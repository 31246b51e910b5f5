"""BASELINE.json's full-size configurations on one GPU, checked through size-independent
properties (tools/fullsize_check.py: counts, bincount(.parse) == .occ, .sai/.last against the
text, every phrase end a trigger and every trigger of a 64 MB slice a phrase end, sampled
dictionary order, unparse of a 200 000-phrase prefix).  The oracle covers the small sizes."""
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
    for (w, p) in [(10, 100), (6, 1000), (32, 50), (16, 500)]:
        assert fc.check_case(sc, text, w, p, f"pangenome 4 GB w{w} p{p}"), f"w={w} p={p}"


def test_config4_random_2gb(pkg, sc):
    """config 4's shape (uniform random ACGT, nearly every phrase distinct) at 2 GB."""
    import fullsize_check as fc
    _need(40)
    text = pkg.synth.random_dna(2_000_000_000, 4, device="cuda")
    assert fc.check_case(sc, text, 10, 100, "random ACGT 2 GB")
