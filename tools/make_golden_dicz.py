#!/usr/bin/env python3
"""Generate tests/golden/golden_dicz.npz: `.dicz` files written by the UNMODIFIED reference
scanner with -c (newscan.cpp:410-413), for the cases listed below (inputs are those of
golden.npz, or seeded).  Pins the -c mode of the oracle restatement and of the CUDA path.

Usage (build container only; needs oracle/_ref):  python tools/make_golden_dicz.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pfp_oracle as orc  # noqa: E402

CASES = ["dna20k_w10_p100", "dna20k_w4_p10", "dna20k_w32_p1000", "first_window", "n11", "n0",
         "all_A", "bytes_3_255", "ends_on_trigger", "pangenome_plain", "pangenome_fasta", "fasta_crlf"]


def main():
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle ref"
    with open(os.path.join(ROOT, "tests", "golden", "cases.json")) as f:
        meta = {c["name"]: c for c in json.load(f)["cases"]}
    npz = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    arrays, out_meta = {}, []
    for name in CASES:
        c = meta[name]
        data = npz[name + "/input"].tobytes()
        tmp = tempfile.mkdtemp(prefix="pfpdicz_")
        try:
            path = os.path.join(tmp, "in.fa" if c["fasta"] else "in.txt")
            with open(path, "wb") as f:
                f.write(data)
            cmd = [orc.ref_exe("newscanNT.x"), path, "-w", str(c["w"]), "-p", str(c["p"]), "-c"]
            if c["fasta"]:
                cmd.append("-f")
            subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            assert not os.path.exists(path + ".dict")
            with open(path + ".dicz", "rb") as f:
                dicz = f.read()
            with open(path + ".parse", "rb") as f:
                assert f.read() == npz[name + "/parse"].tobytes()      # -c leaves the parse alone
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        arrays[name + "/dicz"] = np.frombuffer(dicz, np.uint8)
        out_meta.append(dict(name=name, w=c["w"], p=c["p"], fasta=c["fasta"], dicz_bytes=len(dicz)))
        print(f"{name:28s} dicz {len(dicz)} bytes")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_dicz.npz"), **arrays)
    with open(os.path.join(ROOT, "tests", "golden", "cases_dicz.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden_dicz.py",
                       reference="alshai/Big-BWT newscanNT.x -c [-f] (oracle/_ref, unmodified)",
                       cases=out_meta), f, indent=1)


if __name__ == "__main__":
    main()
