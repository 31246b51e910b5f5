"""The file path of the scanner (csrc/pfp_ingest.cu): K0 -- FASTA extraction on the device
(kseq.h:177-218, newscan.cpp:338-349) -- against the host reader and the reference's golden
cases, and pfpb200_parse_file streaming the input through the pinned ring and the outputs back
(newscan.cpp:332-374, utils.c:33-54) against the in-memory entry points."""
import os
import shutil
import tempfile

import numpy as np
import pytest
import torch

from conftest import assert_same_files, golden, golden_names
from oracle import pfp_oracle as orc

pytestmark = pytest.mark.gpu

FASTA_CASES = [n for n in golden_names() if n.startswith(("fasta", "fastq", "pangenome_fasta"))]
# plain multi-line FASTA: K0 must take these itself (the others may fall back to the host reader)
MUST_BE_ON_DEVICE = {"pangenome_fasta", "pangenome_fasta_w6_p50", "fasta_lower_N", "fasta_empty_record",
                     "fasta_no_trailing_nl", "fasta_header_only", "fasta_gt_inside_line", "fasta_high_byte"}


@pytest.fixture(scope="module")
def sc(pkg):
    s = pkg.pfp.Scanner(0)
    yield s
    s.close()


def k0(sc, data: bytes):
    dev = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).cuda() if data else torch.empty(0, dtype=torch.uint8, device="cuda")
    return sc.fasta_extract_device(dev)


@pytest.mark.parametrize("name", FASTA_CASES)
def test_k0_golden_cases(pkg, sc, name):
    c = golden().case(name)
    want, trunc = pkg.pfp.fasta_extract(c["input"])
    got, supported = k0(sc, c["input"])
    if name in MUST_BE_ON_DEVICE:
        assert supported, f"{name}: plain FASTA must not need the host reader"
    if supported:
        assert not trunc and got == want, f"{name}: device extraction differs from the host reader"


def test_k0_lines_headers_and_records_across_tiles(pkg, sc):
    """Sequence lines and header lines longer than a 16 KB tile, empty lines, records without
    sequence, lower case, no trailing newline; every boundary inside a tile and at tile borders."""
    rng = np.random.default_rng(181)
    parts = []
    for k in range(40):
        hdr = b">rec%d " % k + bytes(rng.integers(33, 127, int(rng.choice([0, 5, 70, 20_000]))).astype(np.uint8))
        hdr = hdr.replace(b"\n", b" ")
        parts.append(hdr + b"\n")
        n = int(rng.choice([0, 1, 59, 60, 61, 5_000, 70_000]))
        seq = pkg.synth.random_dna(n, 300 + k).numpy().tobytes() if n else b""
        if k % 3 == 0:
            seq = seq.lower()
        width = int(rng.choice([60, 70, 16_384, 1 << 20]))
        for o in range(0, len(seq), width):
            parts.append(seq[o:o + width] + b"\n")
        if k % 5 == 0:
            parts.append(b"\n\n")
    data = b"".join(parts)
    for tail in (data, data.rstrip(b"\n"), data + b">last"):
        want, trunc = pkg.pfp.fasta_extract(tail)
        got, supported = k0(sc, tail)
        assert supported and not trunc and got == want
        assert want == orc.fasta_extract(tail)[0]
    # every alignment of the output inside 16 bytes: shift the first record's length
    for extra in range(1, 17):
        d2 = b">a\n" + b"ACGT" * 1000 + b"A" * extra + b"\n" + data
        got, supported = k0(sc, d2)
        assert supported and got == pkg.pfp.fasta_extract(d2)[0]


def test_k0_reports_what_it_does_not_handle(pkg, sc):
    s = pkg.synth.random_dna(500, 182).numpy().tobytes()
    for data in (b"junk\n>x\n" + s + b"\n",                      # junk in front of the first '>'
                 b">x\r\n" + s + b"\r\n",                         # CRLF
                 b"@r\n" + s + b"\n+\n" + b"I" * 500 + b"\n",     # FASTQ
                 b">x\n" + s + b"\n@r2\n" + s + b"\n+\n" + b"I" * 500 + b"\n",
                 b">x\n" + s[:100] + b"\x01" + s[100:] + b"\n",   # invalid byte: the host reader cuts there
                 b">x\n" + s[:100] + b"\xff" + s[100:] + b"\n"):
        got, supported = k0(sc, data)
        assert not supported and got is None
    assert k0(sc, b"") == (b"", True)


def _write(path, data):
    with open(path, "wb") as f:
        f.write(data)


@pytest.mark.parametrize("fasta", [False, True])
def test_parse_file_streams_large_inputs(pkg, fasta):
    """72 MB through pfpb200_parse_file: more chunks than the ring has slots, in both directions;
    the files must be what the in-memory entry point returns for the same text."""
    recs = [r.numpy() for r in pkg.synth.pangenome_records(6_000_000, 12, 183)]
    text = np.concatenate(recs)
    tmp = tempfile.mkdtemp(prefix="pfpfile_")
    sc = pkg.pfp.Scanner(0)
    try:
        path = os.path.join(tmp, "big.fa" if fasta else "big.txt")
        if fasta:
            with open(path, "wb") as f:
                for k, r in enumerate(recs):
                    pkg.synth.to_fasta_np(r, f"hap{k}").tofile(f)
        else:
            text.tofile(path)
        st = sc.parse_file(path, 10, 100, sai=True, fasta=fasta)
        assert st["n_text"] == text.size
        got = orc.collect_files(path)
        want = sc.parse_host(text, 10, 100, sai=True)
        assert_same_files(got, want, f"parse_file fasta={fasta}")
        # segmented .last/.sai, compressed dictionary
        sc.parse_file(path, 10, 100, sai=True, fasta=fasta, nseg=3, compress=True)
        seg = orc.collect_files(path, nseg=3)
        assert seg.last == want.last and seg.sai == want.sai
        with open(path + ".dicz", "rb") as f:
            assert f.read() == orc.dicz_of(want.dict, 10)
    finally:
        sc.close()
        shutil.rmtree(tmp, ignore_errors=True)


@pytest.mark.parametrize("name", FASTA_CASES + ["dna20k_w10_p100", "n0", "n3_lt_w", "invalid_byte", "invalid_byte_at_0",
                                                "bytes_3_255"])
def test_parse_file_golden(pkg, sc, name):
    """Every golden case through the file entry point (device K0 or host reader, whichever the
    bytes need): the five files the unmodified newscanNT.x wrote."""
    c = golden().case(name)
    tmp = tempfile.mkdtemp(prefix="pfpgold_")
    try:
        path = os.path.join(tmp, "in")
        _write(path, c["input"])
        sc.parse_file(path, c["w"], c["p"], sai=True, fasta=c["fasta"])
        assert_same_files(orc.collect_files(path), c, name)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_parse_file_errors(pkg, sc):
    with pytest.raises(pkg.pfp.PfpError) as e:
        sc.parse_file("/nonexistent/file", 10, 100)
    assert e.value.code == -2
