#!/usr/bin/env python3
"""sha256 of the five files the UNMODIFIED reference scanner (oracle/_ref/newscanNT.x -s) writes for
the BASELINE.json configurations at FULL size: config 2 (100 haplotypes x 40 Mbp, w=10 p=100),
corners of the config 5 sweep on the same text, config 4 (8 GB uniform random ACGT).  The texts
come from big-bwt_b200/synth.py on the CPU (a counter-based generator: the GPU emits the same
bytes; the digest of the text itself is stored to prove it).  tools/fullsize_check.py compares the
digests of the CUDA path's outputs with these on the GPU box.

Build container only (needs oracle/_ref, ~60 GB of RAM and scratch disk, minutes to an hour of CPU):
    python tools/make_fullsize_digests.py [--scratch /tmp/fullsize] [--cases config2,sweep,config4]
Writes tests/golden/fullsize_sha256.json (merging with what is there).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pfp_oracle as orc  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize_sha256.json")


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scratch", default="/tmp/fullsize")
    ap.add_argument("--cases", default="config2,sweep,config4")
    a = ap.parse_args()
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle ref"
    synth = load_package().synth
    os.makedirs(a.scratch, exist_ok=True)
    want = a.cases.split(",")
    jobs = []          # (case name, text file, w, p)
    texts = {}
    if "config2" in want or "sweep" in want:
        path = os.path.join(a.scratch, "pangenome_100x40M.txt")
        if not os.path.exists(path):
            t0 = time.time()
            with open(path + ".tmp", "wb") as f:
                for rec in synth.pangenome_records(40_000_000, 100, 2):
                    rec.numpy().tofile(f)
            os.rename(path + ".tmp", path)
            print(f"generated {path} in {time.time() - t0:.0f} s", flush=True)
        texts["pangenome"] = path
        if "config2" in want:
            jobs.append(("config2 w10 p100", path, 10, 100))
        if "sweep" in want:
            jobs += [("sweep w6 p50", path, 6, 50), ("sweep w16 p500", path, 16, 500), ("sweep w32 p1000", path, 32, 1000)]
    if "config4" in want:
        path = os.path.join(a.scratch, "random_8G.txt")
        if not os.path.exists(path):
            t0 = time.time()
            with open(path + ".tmp", "wb") as f:
                step = 1 << 28
                for o in range(0, 8_000_000_000, step):
                    synth.random_dna(min(step, 8_000_000_000 - o), 4, start=o).numpy().tofile(f)
            os.rename(path + ".tmp", path)
            print(f"generated {path} in {time.time() - t0:.0f} s", flush=True)
        texts["random"] = path
        jobs.append(("config4 random 8 GB w10 p100", path, 10, 100))
    # one reference process per case, all at once (each is single-threaded); every case works on
    # its own hard link so that the output files do not collide
    procs = []
    for k, (name, path, w, p) in enumerate(jobs):
        link = os.path.join(a.scratch, f"case{k}_{os.path.basename(path)}")
        if os.path.exists(link):
            os.remove(link)
        os.link(path, link)
        cmd = [orc.ref_exe("newscanNT.x"), link, "-w", str(w), "-p", str(p), "-s", "-P"]
        procs.append((name, link, w, p, time.time(), subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)))
    res = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            res = json.load(f)
    res.setdefault("generator", "tools/make_fullsize_digests.py")
    res.setdefault("reference", "alshai/Big-BWT newscanNT.x -s -P (oracle/_ref, unmodified; -P: probing instead of "
                                "aborting on a 55-bit hash collision, outputs unchanged)")
    res.setdefault("texts", {})
    res.setdefault("cases", {})
    for key, path in texts.items():
        res["texts"][key] = {"bytes": os.path.getsize(path), "sha256": sha_file(path)}
    for name, link, w, p, t0, pr in procs:
        rc = pr.wait()
        if rc != 0:
            print(f"{name}: reference failed with status {rc}", flush=True)
            continue
        res["cases"][name] = {"w": w, "p": p, "reference_seconds": round(time.time() - t0),
                              "sha256": {e: sha_file(f"{link}.{e}") for e in ("dict", "occ", "parse", "last", "sai")},
                              "bytes": {e: os.path.getsize(f"{link}.{e}") for e in ("dict", "occ", "parse", "last", "sai")}}
        print(name, res["cases"][name], flush=True)
        for e in ("dict", "occ", "parse", "last", "sai", "parse_old"):
            try:
                os.remove(f"{link}.{e}")
            except OSError:
                pass
        os.remove(link)
        if os.path.exists(OUT):                 # another instance may have written in the meantime
            with open(OUT) as f:
                disk = json.load(f)
            for k in ("texts", "cases"):
                merged = dict(disk.get(k, {}))
                merged.update(res[k])
                res[k] = merged
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
