#!/usr/bin/env python3
"""Full-size property checks of BASELINE.json configs 2, 4 and 5 on one GPU.

At these sizes the CPU oracle would take minutes to hours, so parity is checked through
size-independent properties of the five output streams (all computed on the device, plus a
host reconstruction of a text prefix from .dict + .parse):
  * sum(.occ) == #phrases, bincount(.parse) == .occ, every rank in 1..d
  * .sai strictly increasing, last value n + w; .last[j] == T[sai[j] - w - 1]
  * every phrase end is a trigger: T[sai-w .. sai) hashes to 0 mod p (re-computed with torch)
  * .dict: d words, EVERY adjacent pair strictly increasing (device kernel), 0x01 terminators, final 0x00
  * unparse of the first 200 000 phrases == the text prefix (host), and the FULL round trip on the device:
    pfpb200_unparse_device(.dict, .parse) == text, every byte
and, where tests/golden/fullsize_sha256.json holds them (tools/make_fullsize_digests.py ran the
unmodified newscanNT.x on the same full-size text in the build container), BYTE-EXACT parity:
sha256 of each of the five streams == the reference's.
usage: fullsize_check.py [--config pangenome|random|sweep|all] [--small]
Prints one JSON line per case; exit status 1 on any failed property.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

PW = 1999999973


class DevView:
    """torch view of a raw device pointer (no copy) through __cuda_array_interface__."""
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def view(ptr, n, typestr="|u1"):
    return torch.as_tensor(DevView(ptr, n, typestr), device="cuda")


def window_hash_mod_p(text, ends, w, p):
    """(KR hash of T[e-w+1..e]) % p for int64 tensor of end positions (newscan.cpp:194-202)."""
    h = torch.zeros_like(ends)
    for i in range(w):
        c = text[ends - w + 1 + i].to(torch.int64)
        h = (h * 256 + c) % PW
    return h % p


def load_digests():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize_sha256.json")) as f:
            return json.load(f)
    except OSError:
        return {"cases": {}, "texts": {}}


def sha_dev(ptr, nbytes):
    """sha256 of a device buffer (copied to the host in 256 MB pieces)."""
    import hashlib
    h = hashlib.sha256()
    step = 1 << 28
    for o in range(0, nbytes, step):
        h.update(view(ptr + o, min(step, nbytes - o)).cpu().numpy().tobytes())
    return h.hexdigest()


def check_case(sc, text, w, p, name, digest_key=None, text_key=None, prefix_phrases=200_000, pipeline=False):
    n = text.numel()
    sc.parse_device(text, w, p, sai=True)               # warm-up: arena growth, table sizing hint
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sc.parse_device(text, w, p, sai=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = sc.stats.as_dict()
    P, d = out.n_phrases, out.n_distinct
    res = {"case": name, "n": n, "w": w, "p": p, "phrases": P, "distinct": d, "dict_bytes": out.dict_bytes,
           "ms_gpu": round(st["ms_total"], 3), "GBps": round(n / st["ms_total"] / 1e6, 1), "wall_s": round(dt, 3),
           "stages_ms": {k[3:]: round(v, 2) for k, v in st.items() if k.startswith("ms_") and k not in ("ms_total", "ms_h2d", "ms_d2h")},
           "rank_rounds": st["rank_rounds"]}
    fails = []
    parse = view(out.parse, P, "<u4").to(torch.int64)
    occ = view(out.occ, d, "<u4").to(torch.int64)
    last = view(out.last, P)
    sai = view(out.sai, 5 * P).view(P, 5).to(torch.int64)
    pos = sai[:, 0] | (sai[:, 1] << 8) | (sai[:, 2] << 16) | (sai[:, 3] << 24) | (sai[:, 4] << 32)
    if int(occ.sum()) != P: fails.append("sum(occ) != phrases")
    if int(parse.min()) < 1 or int(parse.max()) > d: fails.append("rank out of range")
    if not torch.equal(torch.bincount(parse, minlength=d + 1)[1:], occ): fails.append("bincount(parse) != occ")
    if P > 1 and not bool((pos[1:] > pos[:-1]).all()): fails.append(".sai not increasing")
    if int(pos[-1]) != n + w: fails.append(".sai last != n + w")
    e = pos[:-1] - 1                                     # trigger positions
    if e.numel():
        lidx = e - w
        ok = torch.where(lidx >= 0, text[lidx.clamp(min=0)], torch.full_like(last[:-1], 2)) == last[:-1]
        if not bool(ok.all()): fails.append(".last mismatch")
        if int(last[-1]) != int(text[n - 1]): fails.append(".last final mismatch")
        step = max(1, e.numel() // 2_000_000)             # sampled trigger re-computation
        if int(window_hash_mod_p(text, e[::step], w, p).abs().sum()) != 0: fails.append("phrase end is not a trigger")
    # trigger count cross-check on a 64 MB slice: every trigger in the slice must be a phrase end
    m = min(n, 1 << 26)
    allpos = torch.arange(w - 1, m, dtype=torch.int64, device=text.device)
    trig = allpos[window_hash_mod_p(text, allpos, w, p) == 0]
    mine = e[e < m]
    if not torch.equal(trig, mine): fails.append("trigger set differs on the first 64 MB")
    # dictionary (host): terminators, order of sampled adjacent pairs, prefix reconstruction
    dic = view(out.dict, out.dict_bytes).cpu().numpy()
    if dic[-1] != 0: fails.append(".dict does not end with 0x00")
    seps = np.flatnonzero(dic == 1)
    if seps.size != d: fails.append("#0x01 != distinct")
    starts = np.concatenate([[0], seps[:-1] + 1])
    dd = view(out.dict, out.dict_bytes)
    seps_dev = torch.nonzero(dd == 1).flatten()
    torch.cuda.synchronize()                             # the library runs on its own stream
    bad = sc.check_dict_order(dd, seps_dev) if seps_dev.numel() == d else -1
    res["dict_pairs_checked"] = max(d - 1, 0)
    if bad: fails.append(f".dict order violated in {bad} adjacent pairs")
    del seps_dev
    k = min(P, prefix_phrases)
    ranks = parse[:k].cpu().numpy()
    parts = []
    for j, r in enumerate(ranks):
        wd = dic[starts[r - 1]:seps[r - 1]]
        parts.append(wd if j == 0 else wd[w:])
    rec = np.concatenate(parts)[1:]
    if k == P: rec = rec[:len(rec) - w]
    if not np.array_equal(rec, text[:rec.size].cpu().numpy()): fails.append("unparse(prefix) != text prefix")
    # byte-exact parity with the reference at full size, through digests
    dg = load_digests()
    if text_key and text_key in dg["texts"]:
        t_sha = sha_dev(text.data_ptr(), n)
        res["text_sha256_matches_cpu_generator"] = t_sha == dg["texts"][text_key]["sha256"]
        if not res["text_sha256_matches_cpu_generator"]: fails.append("the GPU generator's text differs from the CPU generator's")
    if digest_key and digest_key in dg["cases"]:
        want = dg["cases"][digest_key]["sha256"]
        sizes = {"dict": out.dict_bytes, "occ": 4 * d, "parse": 4 * P, "last": P, "sai": 5 * P}
        ptrs = {"dict": out.dict, "occ": out.occ, "parse": out.parse, "last": out.last, "sai": out.sai}
        got = {e: sha_dev(ptrs[e], sizes[e]) for e in want}
        res["sha256_vs_reference"] = {e: got[e] == want[e] for e in want}
        res["sha256"] = {e: got[e][:16] for e in got}
        for e in want:
            if got[e] != want[e]: fails.append(f".{e} sha256 differs from newscanNT.x")
    else:
        res["sha256_vs_reference"] = None
    # the round trip at full size, every byte, on the device: unparse(.dict, .parse) == text
    # (pfpb200_unparse_device keeps the outputs of the parse alive; reference: unparse.c)
    del parse, occ, last, sai, pos
    torch.cuda.empty_cache()
    try:
        ptr, nt, ms_up = sc.unparse_device(out.dict, out.dict_bytes, out.parse, P, strip_w=w)
        torch.cuda.synchronize()
        same = nt == n and bool(torch.equal(view(ptr, nt), text))
        res["unparse_round_trip"] = {"bytes": nt, "ms": round(ms_up, 3), "ok": same}
        if not same: fails.append("unparse(.dict, .parse) != text")
    except Exception as e:  # noqa: BLE001
        res["unparse_round_trip"] = {"error": str(e)[:200]}
        fails.append("unparse failed")
    # the stages after the parse at full size, chained in HBM on the outputs above: bwtparse, pfbwt -S;
    # digests of .ilist .bwlast .bwsai .bwt .sa against the unmodified reference chain's
    # (tools/make_fullsize_digests_pipeline.py), where the build container has produced them
    if pipeline:
        try:
            # (a context of their own, closed afterwards: the scratch of the suffix sorts -- 58 B per dictionary
            # byte -- goes back to the device instead of staying in the parse context's arena)
            sc2 = type(sc)(sc.device)
            bp = sc2.bwtparse_device(out.parse, out.n_phrases, out.last, out.sai)
            r = sc2.pfbwt_device(out.dict, out.dict_bytes, out.occ, out.n_distinct, bp.ilist, bp.bwlast, bp.bwsai,
                                 bp.n_out, w, 1)      # PFPB200_PFBWT_SA
            res["pipeline"] = {"ms_bwtparse": round(bp.ms_total, 2), "ms_pfbwt": round(r.ms_total, 2),
                               "bwt_bytes": r.n_bwt, "easy": r.easy, "hard": r.hard}
            if r.n_bwt != n + 1: fails.append("pfbwt: |BWT| != n + 1")
            ref = dg["cases"].get(digest_key or "", {}).get("pipeline")
            if ref:
                got = {"ilist": sha_dev(bp.ilist, 4 * bp.n_out), "bwlast": sha_dev(bp.bwlast, bp.n_out),
                       "bwsai": sha_dev(bp.bwsai, 5 * bp.n_out), "bwt": sha_dev(r.bwt, r.n_bwt),
                       "sa": sha_dev(r.sa, 5 * r.n_sa)}
                res["pipeline"]["sha256_vs_reference"] = {e: got[e] == ref["sha256"][e] for e in got}
                for e in got:
                    if got[e] != ref["sha256"][e]: fails.append(f".{e} sha256 differs from the reference chain")
            else:
                res["pipeline"]["sha256_vs_reference"] = None
            sc2.close()
        except Exception as e:  # noqa: BLE001
            res["pipeline"] = {"error": str(e)[:300]}
            fails.append("pipeline failed")
    res["ok"] = not fails
    res["fails"] = fails
    print(json.dumps(res), flush=True)
    return not fails


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all", choices=["pangenome", "random", "sweep", "all"])
    ap.add_argument("--small", action="store_true", help="1/16 of the sizes (smoke run)")
    a = ap.parse_args()
    from __graft_entry__ import load_package
    pkg = load_package()
    sc = pkg.pfp.Scanner(0)
    scale = 16 if a.small else 1
    ok = True
    if a.config in ("pangenome", "sweep", "all"):
        text = pkg.synth.pangenome_text(40_000_000 // scale, 100, 2, device="cuda")
        if a.config in ("pangenome", "all"):
            ok &= check_case(sc, text, 10, 100, "config2 pangenome 100 x 40 Mbp",
                             digest_key=None if a.small else "config2 w10 p100", text_key=None if a.small else "pangenome",
                             pipeline=True)
        if a.config in ("sweep", "all"):
            for w in (6, 10, 16, 32):
                for p in (50, 100, 500, 1000):
                    ok &= check_case(sc, text, w, p, "config5 sweep", digest_key=None if a.small else f"sweep w{w} p{p}")
        del text
        torch.cuda.empty_cache()
    if a.config in ("random", "all"):
        text = pkg.synth.random_dna(8_000_000_000 // scale, 4, device="cuda")
        ok &= check_case(sc, text, 10, 100, "config4 random ACGT 8 GB",
                         digest_key=None if a.small else "config4 random 8 GB w10 p100", text_key=None if a.small else "random")
    sc.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
