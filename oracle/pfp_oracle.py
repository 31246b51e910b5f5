"""ctypes face of oracle/libpfp_oracle.so and runner for the reference binaries in oracle/_ref.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under big-bwt_b200/ may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpfp_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

THREADED_RULE = 1


class _Result(C.Structure):
    _fields_ = [
        ("n_text", C.c_uint64), ("n_phrases", C.c_uint64), ("n_distinct", C.c_uint64),
        ("sum_word_len", C.c_uint64), ("dict_len", C.c_uint64),
        ("dict", C.POINTER(C.c_uint8)), ("occ", C.POINTER(C.c_uint32)),
        ("parse", C.POINTER(C.c_uint32)), ("last", C.POINTER(C.c_uint8)),
        ("sai", C.POINTER(C.c_uint8)),
        ("sec_scan", C.c_double), ("sec_sort", C.c_double), ("sec_remap", C.c_double),
    ]


@dataclass
class PfpFiles:
    """The five output files of the parsing stage, as bytes."""
    dict: bytes
    occ: bytes
    parse: bytes
    last: bytes
    sai: bytes
    n_text: int = 0
    n_phrases: int = 0
    n_distinct: int = 0
    sum_word_len: int = 0
    seconds: tuple = (0.0, 0.0, 0.0)

    def names(self):
        return ("dict", "occ", "parse", "last", "sai")


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.pfp_oracle_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.POINTER(_Result)]
        L.pfp_oracle_parse.restype = C.c_int
        L.pfp_oracle_free.argtypes = [C.POINTER(_Result)]
        L.pfp_oracle_triggers.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_uint64]
        L.pfp_oracle_triggers.restype = C.c_uint64
        L.pfp_oracle_fasta_extract.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p,
                                               C.POINTER(C.c_int)]
        L.pfp_oracle_fasta_extract.restype = C.c_uint64
        L.pfp_oracle_kr_hash.argtypes = [C.c_void_p, C.c_uint64]
        L.pfp_oracle_kr_hash.restype = C.c_uint64
        L.pfp_oracle_window_hash.argtypes = [C.c_void_p, C.c_uint32]
        L.pfp_oracle_window_hash.restype = C.c_uint64
        _lib = L
    return _lib


def _as_u8(text) -> np.ndarray:
    if isinstance(text, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(text), dtype=np.uint8)
    a = np.ascontiguousarray(text)
    assert a.dtype == np.uint8
    return a


def parse(text, w=10, p=100, flags=0) -> PfpFiles:
    a = _as_u8(text)
    r = _Result()
    rc = lib().pfp_oracle_parse(a.ctypes.data if a.size else None, a.size, w, p, flags, C.byref(r))
    if rc != 0:
        raise RuntimeError(f"pfp_oracle_parse failed: {rc}")
    try:
        np_, nd = r.n_phrases, r.n_distinct
        out = PfpFiles(
            dict=C.string_at(r.dict, r.dict_len),
            occ=C.string_at(r.occ, 4 * nd),
            parse=C.string_at(r.parse, 4 * np_),
            last=C.string_at(r.last, np_),
            sai=C.string_at(r.sai, 5 * np_),
            n_text=r.n_text, n_phrases=np_, n_distinct=nd, sum_word_len=r.sum_word_len,
            seconds=(r.sec_scan, r.sec_sort, r.sec_remap),
        )
    finally:
        lib().pfp_oracle_free(C.byref(r))
    return out


def triggers(text, w=10, p=100, flags=0) -> np.ndarray:
    a = _as_u8(text)
    cap = a.size + 1
    out = np.empty(cap, dtype=np.uint64)
    k = lib().pfp_oracle_triggers(a.ctypes.data if a.size else None, a.size, w, p, flags,
                                  out.ctypes.data, cap)
    return out[:k].copy()


def fasta_extract(file_bytes):
    a = _as_u8(file_bytes)
    out = np.empty(a.size + 1, dtype=np.uint8)
    tr = C.c_int(0)
    n = lib().pfp_oracle_fasta_extract(a.ctypes.data if a.size else None, a.size,
                                       out.ctypes.data, C.byref(tr))
    return out[:n].tobytes(), bool(tr.value)


def dicz_of(dict_bytes: bytes, w: int) -> bytes:
    """The `.dicz` file newscan -c writes instead of `.dict` (newscan.cpp:410-413): every word
    loses its last w bytes and, if it starts with Dollar (0x02), that byte too."""
    assert dict_bytes[-1:] == b"\x00"
    out = []
    for wd in dict_bytes[:-1].split(b"\x01")[:-1]:
        assert len(wd) > w
        wd = wd[:-w]
        out.append((wd[1:] if wd[:1] == b"\x02" else wd) + b"\x01")
    return b"".join(out) + b"\x00"


def kr_hash(s: bytes) -> int:
    a = _as_u8(s)
    return lib().pfp_oracle_kr_hash(a.ctypes.data if a.size else None, a.size)


def window_hash(s: bytes) -> int:
    a = _as_u8(s)
    return lib().pfp_oracle_window_hash(a.ctypes.data, a.size)


# ---------------------------------------------------------------------------------------
# the unmodified reference, compiled into oracle/_ref by oracle/Makefile
# ---------------------------------------------------------------------------------------
def have_ref(exe="newscanNT.x") -> bool:
    return os.path.exists(os.path.join(REF_DIR, exe))


def ref_exe(name: str) -> str:
    return os.path.join(REF_DIR, name)


def _read(path):
    with open(path, "rb") as f:
        return f.read()


def collect_files(base: str, nseg: int = 0, sai: bool = True) -> PfpFiles:
    """Read <base>.dict/.occ/.parse and the (possibly segmented) .last/.sai streams."""
    def stream(ext):
        if nseg == 0:
            return _read(f"{base}.{ext}")
        return b"".join(_read(f"{base}.{i}.{ext}") for i in range(nseg))
    d = _read(base + ".dict")
    occ = _read(base + ".occ")
    parse_ = _read(base + ".parse")
    return PfpFiles(dict=d, occ=occ, parse=parse_, last=stream("last"),
                    sai=stream("sai") if sai else b"",
                    n_phrases=len(parse_) // 4, n_distinct=len(occ) // 4)


def run_reference(data: bytes, w=10, p=100, fasta=False, exe="newscanNT.x", threads=0,
                  sai=True, keep_dir=None, timeout=600) -> PfpFiles:
    """Run a reference scanner on `data` written to a scratch file; return its outputs."""
    tmp = keep_dir or tempfile.mkdtemp(prefix="pfpref_")
    try:
        path = os.path.join(tmp, "in.fa" if fasta else "in.txt")
        with open(path, "wb") as f:
            f.write(data)
        cmd = [ref_exe(exe), path, "-w", str(w), "-p", str(p)]
        if sai:
            cmd.append("-s")
        if fasta:
            cmd.append("-f")
        if threads:
            cmd += ["-t", str(threads)]
        subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=timeout)
        return collect_files(path, nseg=threads, sai=sai)
    finally:
        if keep_dir is None:
            shutil.rmtree(tmp, ignore_errors=True)
