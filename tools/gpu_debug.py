#!/usr/bin/env python3
"""Stage-by-stage diagnosis on a GPU box: prints where the CUDA path first departs from the oracle."""
import sys, os, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
from oracle import pfp_oracle as orc
import torch
pkg = load_package()
sc = pkg.pfp.Scanner(0)
print("scanner ok", torch.cuda.get_device_name(0))
for (n, w, p) in [(1000, 10, 100), (100000, 10, 100), (300000, 4, 10), (300000, 19, 64), (2000000, 10, 100)]:
    text = pkg.synth.random_dna(n, 21).numpy()
    try:
        got, ms = sc.scan_triggers(torch.from_numpy(text).cuda(), w, p)
        want = orc.triggers(text.tobytes(), w, p)
        ok = np.array_equal(got, want)
        print(f"scan n={n} w={w} p={p}: {len(got)} vs {len(want)} ok={ok} ms={ms:.3f}")
        if not ok:
            m = min(len(got), len(want)); bad = np.nonzero(got[:m] != want[:m])[0]
            print("  first diff idx", bad[:5], got[:8], want[:8])
    except Exception:
        traceback.print_exc()
for (n, w, p) in [(1000, 10, 100), (40000, 10, 100), (1000000, 10, 100)]:
    text = pkg.synth.pangenome_text(n // 10, 10, 3).numpy().tobytes()
    try:
        got = sc.parse_host(text, w, p)
        want = orc.parse(text, w, p)
        print(f"parse n={len(text)}: phrases {got.n_phrases}/{want.n_phrases} distinct {got.n_distinct}/{want.n_distinct}",
              {e: getattr(got, e) == getattr(want, e) for e in ("dict", "occ", "parse", "last", "sai")})
        print("  stats", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in got.stats.items()})
    except Exception:
        traceback.print_exc()
