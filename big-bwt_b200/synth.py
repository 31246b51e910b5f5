"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything is a counter-based function of (seed, haplotype, site) built from the
splitmix64 finaliser in wrapping int64 torch arithmetic, so the same bytes come out on the
CPU (tests, oracle, reference arm) and on the GPU (bench), for any prefix of the input.

  base genome     : uniform ACGT
  haplotype h     : the base with independent per-site mutations, probability 1049/2^20
                    (0.1 %): 80 % SNP to a different base, 10 % 1-bp insertion after the
                    site, 10 % deletion of the site
  pan-genome text : the haplotypes concatenated with no separator, i.e. exactly the text the
                    reference extracts from a one-record-per-haplotype FASTA with `-f`
"""
from __future__ import annotations

import numpy as np
import torch

_M63 = (1 << 63) - 1


def _s64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >> 63 else v


_C0 = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)
MUT_THRESHOLD = 1049  # of 2^20


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 output function on an int64 tensor (wrapping arithmetic)."""
    x = x + _C0
    x = (x ^ _lsr(x, 30)) * _C1
    x = (x ^ _lsr(x, 27)) * _C2
    return x ^ _lsr(x, 31)


def _mix_int(v: int) -> int:
    m = (1 << 64) - 1
    v = (v + 0x9E3779B97F4A7C15) & m
    v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & m
    v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & m
    return v ^ (v >> 31)


_ACGT = (65, 67, 71, 84)


def _codes_to_ascii(codes: torch.Tensor) -> torch.Tensor:
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=codes.device)
    return lut[codes]


def base_codes(n: int, seed: int, device="cpu", start: int = 0) -> torch.Tensor:
    """2-bit codes of sites [start, start+n) of the base genome of `seed`."""
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    key = _s64(_mix_int(seed))
    return _lsr(_mix(i ^ key), 33) & 3


def random_dna(n: int, seed: int, device="cpu", start: int = 0) -> torch.Tensor:
    """Uniform random ACGT text (config 4 shape), bytes [start, start+n) of stream `seed`."""
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 24
    for o in range(0, n, step):
        m = min(step, n - o)
        out[o:o + m] = _codes_to_ascii(base_codes(m, seed, device, start + o))
    return out


def haplotype(codes: torch.Tensor, seed: int, h: int, site0: int = 0) -> torch.Tensor:
    """ASCII sequence of haplotype h given the base codes of sites [site0, site0+len)."""
    dev = codes.device
    n = codes.numel()
    i = torch.arange(site0, site0 + n, dtype=torch.int64, device=dev)
    key = _s64(_mix_int((seed << 20) ^ (h + 1) * 0x9E3779B1))
    u = _mix(i ^ key)
    mut = (u & 0xFFFFF) < MUT_THRESHOLD
    kind = _lsr(u, 20) % 10
    snp = mut & (kind < 8)
    ins = mut & (kind == 8)
    dele = mut & (kind == 9)
    alt = (codes + 1 + (_lsr(u, 32) % 3)) & 3
    emit = torch.where(snp, alt, codes)
    extra = _lsr(u, 40) & 3
    lens = 1 + ins.to(torch.int64) - dele.to(torch.int64)
    ends = torch.cumsum(lens, 0)
    total = int(ends[-1].item()) if n else 0
    offs = ends - lens
    out = torch.empty(total, dtype=torch.uint8, device=dev)
    keep = ~dele
    out[offs[keep]] = _codes_to_ascii(emit[keep])
    out[offs[ins] + 1] = _codes_to_ascii(extra[ins])
    return out


def pangenome_records(base_len: int, n_hap: int, seed: int, device="cpu", first_hap: int = 0):
    """Yield the haplotype sequences (uint8 tensors) one by one."""
    codes = base_codes(base_len, seed, device)
    for h in range(first_hap, first_hap + n_hap):
        yield haplotype(codes, seed, h)


def pangenome_text(base_len: int, n_hap: int, seed: int, device="cpu",
                   first_hap: int = 0) -> torch.Tensor:
    """Concatenated haplotypes = the text T the parser sees for the FASTA of these records.
    Written haplotype by haplotype into one buffer (a 64 GB shard must not exist twice)."""
    if n_hap <= 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    # a haplotype is at most base_len * (1 + insertion rate) long: 0.1 % of the sites mutate, one in
    # ten of those inserts a base; 0.05 % of slack per haplotype is 5 times that
    cap = n_hap * (base_len + base_len // 2000 + 1024)
    out = torch.empty(cap, dtype=torch.uint8, device=device)
    pos = 0
    for rec in pangenome_records(base_len, n_hap, seed, device, first_hap):
        n = rec.numel()
        out[pos:pos + n] = rec
        pos += n
        del rec
    return out[:pos]


# BASELINE.json configs[0] names the reference's bundled yeast.fasta, which is not in the mount
# (SURVEY section 0).  Stand-in of the same shape: the 16 chromosomes + mitochondrion of S. cerevisiae
# S288C by length (12.16 Mbp), i.i.d. bases with 38 % GC.
YEAST_LENGTHS = (230218, 813184, 316620, 1531933, 576874, 270161, 1090940, 562643, 439888, 745751, 666816,
                 1078177, 924431, 784333, 1091291, 948066, 85779)
YEAST_NAMES = tuple("chr" + r for r in ("I II III IV V VI VII VIII IX X XI XII XIII XIV XV XVI".split())) + ("chrM",)


def yeast_like_records(seed: int = 1, device="cpu"):
    """The 17 sequences (uint8 tensors) of the yeast-shaped stand-in."""
    out = []
    for k, n in enumerate(YEAST_LENGTHS):
        i = torch.arange(n, dtype=torch.int64, device=device)
        u = _lsr(_mix(i ^ _s64(_mix_int((seed << 8) ^ (k + 1)))), 44)          # 20 random bits
        # A 31 %, C 19 %, G 19 %, T 31 %
        code = (u >= 325058).to(torch.int64) + (u >= 524288).to(torch.int64) + (u >= 723518).to(torch.int64)
        lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
        out.append(lut[code])
    return out


def to_fasta(records, names=None, width: int = 60, newline: bytes = b"\n") -> bytes:
    """Wrap sequences (uint8 arrays/tensors/bytes) as FASTA records with `width`-column lines."""
    chunks = []
    for k, rec in enumerate(records):
        if isinstance(rec, torch.Tensor):
            rec = rec.cpu().numpy()
        b = bytes(rec) if not isinstance(rec, np.ndarray) else rec.tobytes()
        name = names[k] if names else f"hap{k}"
        chunks.append(b">" + name.encode() + newline)
        for o in range(0, len(b), width):
            chunks.append(b[o:o + width] + newline)
    return b"".join(chunks)


def to_fasta_np(seq: np.ndarray, name: str, width: int = 60) -> np.ndarray:
    """Vectorised single-record FASTA wrapping for large sequences."""
    n = seq.size
    full = n // width
    rem = n - full * width
    body = np.empty(full * (width + 1) + (rem + 1 if rem else 0), dtype=np.uint8)
    if full:
        blk = body[:full * (width + 1)].reshape(full, width + 1)
        blk[:, :width] = seq[:full * width].reshape(full, width)
        blk[:, width] = 10
    if rem:
        body[full * (width + 1):-1] = seq[full * width:]
        body[-1] = 10
    head = np.frombuffer(b">" + name.encode() + b"\n", dtype=np.uint8)
    return np.concatenate([head, body])
