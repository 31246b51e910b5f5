"""The stage-level C ABI of sharded parsing on ONE GPU: G shards are driven in sequence through
G contexts, the collectives of ShardedParser replaced by plain tensor copies, for both merge
modes.  (The real multi-process path: tests/test_shards_cpu.py over gloo, and
tools/multi_gpu_check.py under torchrun on >= 2 GPUs.)"""
import subprocess
import sys
import os

import numpy as np
import pytest
import torch

from conftest import FILES, ROOT
from oracle import pfp_oracle as orc

pytestmark = pytest.mark.gpu


def emulate(pkg, text: bytes, cuts, w, p, mode, halo=64, front=4096):
    sh = pkg.shards
    G = len(cuts) - 1
    dev = torch.device("cuda", 0)
    n = len(text)
    tb = np.frombuffer(text, np.uint8)
    starts, sizes = cuts[:-1], [cuts[k + 1] - cuts[k] for k in range(G)]
    bes = [sh.CudaBackend(0) for _ in range(G)]
    fronts = [min(front, starts[g]) for g in range(G)]
    bufs = []
    for g in range(G):
        b = torch.full((fronts[g] + sizes[g],), 1, dtype=torch.uint8, device=dev)   # poison the front
        b[fronts[g]:] = torch.from_numpy(tb[cuts[g]:cuts[g + 1]].copy()).to(dev)
        bufs.append(b)

    def copy(ops):
        for src, dst, lo, hi in ops:
            a = fronts[src] + lo - starts[src]
            c = fronts[dst] - (starts[dst] - lo)
            bufs[dst][c:c + hi - lo] = bufs[src][a:a + hi - lo]
    copy(sh.transfers([(max(0, starts[q] - halo), starts[q]) for q in range(G)], starts, sizes))
    scans = [bes[g].shard_scan(bufs[g], starts[g] - fronts[g], starts[g], starts[g] + sizes[g], n, g == G - 1,
                               w, p, True) for g in range(G)]
    nts, lasts = [s[0] for s in scans], [s[2] for s in scans]
    copy(sh.transfers(sh.head_requests(starts, sizes, nts, lasts, w, halo), starts, sizes))
    wds = [bes[g].shard_words(sh.first_phrase_start(g, nts, lasts, w)) for g in range(G)]
    if mode == "replicate":
        cat = {k: torch.cat([wd[k] for wd in wds]) for k in ("fpa", "fpb", "len", "count", "uwords", "pool")}
        m = bes[0].dict_merge(cat["fpa"], cat["fpb"], cat["len"], cat["count"], cat["uwords"], cat["pool"], w)
        bases = np.concatenate([[0], np.cumsum([wd["n_words"] for wd in wds])])
        rows = [m["rank_of_entry"][bases[g]:bases[g + 1]].contiguous() for g in range(G)]
        dict_b = m["dict"].cpu().numpy().tobytes()
        occ_b = m["occ"].cpu().numpy().tobytes()
    else:
        samples = np.concatenate([bes[g].first_keys(wds[g]).cpu().numpy().view(np.uint64)[::7] for g in range(G)
                                  if wds[g]["n_words"]])
        vals = np.sort(samples)
        sp = np.array([vals[(q + 1) * vals.size // G] for q in range(G - 1)], dtype=np.uint64)
        rts = [bes[g].route(wds[g], sp, G) for g in range(G)]
        merged, dict_b, occ_b = [], b"", b""
        for o in range(G):                     # owner o receives slice o of every source
            wparts, pparts = [], []
            for src in range(G):
                a = sum(rts[src]["words_to"][:o])
                wparts.append(rts[src]["words"][32 * a:32 * (a + rts[src]["words_to"][o])])
                b = sum(rts[src]["pool_to"][:o])
                pparts.append(rts[src]["pool"][b:b + rts[src]["pool_to"][o]])
            m = bes[o].dict_merge_words(torch.cat(wparts).contiguous(), torch.cat(pparts).contiguous(), w)
            merged.append(m)
            d = m["dict"].cpu().numpy().tobytes()
            dict_b += d if o == G - 1 else d[:-1]
            occ_b += m["occ"].cpu().numpy().tobytes()
        offs = np.concatenate([[0], np.cumsum([m["n_distinct"] for m in merged])])
        rows = []
        for src in range(G):
            back = []
            for o in range(G):
                a = sum(rts[q]["words_to"][o] for q in range(src))
                back.append(merged[o]["rank_of_entry"][a:a + rts[src]["words_to"][o]] + int(offs[o]))
            back = torch.cat(back).to(torch.int32)
            row = torch.empty(wds[src]["n_words"], dtype=torch.int32, device=dev)
            row[rts[src]["perm"].long()] = back
            rows.append(row)
    torch.cuda.synchronize()
    parses = [bes[g].shard_remap(rows[g], wds[g]["n_phrases"]) for g in range(G)]
    out = {"dict": dict_b, "occ": occ_b,
           "parse": b"".join(t.cpu().numpy().tobytes() for t in parses),
           "last": b"".join(wd["last"].cpu().numpy().tobytes() for wd in wds),
           "sai": b"".join(wd["sai"].cpu().numpy().tobytes() for wd in wds)}
    for be in bes:
        be.scanner.close()
    return out


@pytest.mark.parametrize("mode", ["replicate", "partition"])
@pytest.mark.parametrize("G,w,p", [(2, 10, 100), (3, 10, 100), (4, 6, 50), (3, 16, 500)])
def test_shard_stages_match_oracle(pkg, mode, G, w, p):
    text = pkg.synth.pangenome_text(40_000, 12, 17).numpy().tobytes()
    n = len(text)
    cuts = [0] + [n * k // G + 5 * k for k in range(1, G)] + [n]
    got = emulate(pkg, text, cuts, w, p, mode)
    want = orc.parse(text, w, p)
    for ext in FILES:
        assert got[ext] == getattr(want, ext), f".{ext} differs ({mode}, G={G})"


@pytest.mark.parametrize("mode", ["replicate", "partition"])
def test_shard_without_triggers(pkg, mode):
    a = pkg.synth.random_dna(30_000, 3).numpy().tobytes()
    text = a + b"N" * 25_000 + a[:20_000]
    cuts = [0, 32_000, 50_000, len(text)]      # shard 1 lies inside the N run: it owns no phrase
    got = emulate(pkg, text, cuts, 10, 100, mode, halo=64, front=1 << 16)
    want = orc.parse(text, 10, 100)
    for ext in FILES:
        assert got[ext] == getattr(want, ext), f".{ext} differs"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["replicate", "partition"])
def test_two_gpu_torchrun(mode):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tools", "multi_gpu_check.py"), mode],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
