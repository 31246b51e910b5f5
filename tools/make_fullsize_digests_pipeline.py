#!/usr/bin/env python3
"""sha256 of what the UNMODIFIED reference chain writes AFTER the parse for BASELINE config 2 at FULL
size (100 haplotypes x 40 Mbp, w=10 p=100): oracle/_ref/newscanNT.x -s -P, then bwtparse -s
(.ilist .bwlast .bwsai), then pfbwtNT.x -S (.bwt .sa).  The text comes from big-bwt_b200/synth.py
on the CPU (same bytes as the GPU generator; tests/golden/fullsize_sha256.json holds its digest).
tools/fullsize_check.py --pipeline compares the digests of the GPU stages' outputs with these.

Build container only (needs oracle/_ref, ~12 GB of RAM, ~35 GB of scratch disk, ~30 CPU-minutes):
    python tools/make_fullsize_digests_pipeline.py [--scratch /tmp/fullsize]
Merges "pipeline" into the case "config2 w10 p100" of tests/golden/fullsize_sha256.json."""
import argparse, hashlib, json, os, subprocess, sys, time
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
    a = ap.parse_args()
    synth = load_package().synth
    os.makedirs(a.scratch, exist_ok=True)
    path = os.path.join(a.scratch, "pangenome_100x40M.txt")
    if not os.path.exists(path):
        t0 = time.time()
        with open(path + ".tmp", "wb") as f:
            for rec in synth.pangenome_records(40_000_000, 100, 2):
                rec.numpy().tofile(f)
        os.rename(path + ".tmp", path)
        print(f"generated {path} in {time.time() - t0:.0f} s", flush=True)
    base = os.path.join(a.scratch, "pipe_" + os.path.basename(path))
    if os.path.exists(base):
        os.remove(base)
    os.link(path, base)
    secs = {}
    for name, cmd in (("newscanNT.x", [orc.ref_exe("newscanNT.x"), base, "-w", "10", "-p", "100", "-s", "-P"]),
                      ("bwtparse", [orc.ref_exe("bwtparse"), base, "-s"]),
                      ("pfbwtNT.x", [orc.ref_exe("pfbwtNT.x"), "-w", "10", "-S", base])):
        t0 = time.time()
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        secs[name] = round(time.time() - t0)
        print(name, secs[name], "s", flush=True)
    exts = ("ilist", "bwlast", "bwsai", "bwt", "sa")
    pipe = {"reference_seconds": secs, "cmd": "newscanNT.x -s -P; bwtparse -s; pfbwtNT.x -w 10 -S (oracle/_ref, unmodified)",
            "sha256": {e: sha_file(f"{base}.{e}") for e in exts}, "bytes": {e: os.path.getsize(f"{base}.{e}") for e in exts}}
    print(pipe, flush=True)
    with open(OUT) as f:
        res = json.load(f)
    res["cases"].setdefault("config2 w10 p100", {})["pipeline"] = pipe
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
    for e in exts + ("dict", "occ", "parse", "last", "sai", "parse_old"):
        try:
            os.remove(f"{base}.{e}")
        except OSError:
            pass
    os.remove(base)


if __name__ == "__main__":
    main()
