import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
from oracle import pfp_oracle as orc
import test_shards_gpu as T
pkg = load_package()
world = 8
cases = []
cases.append(("pangenome", pkg.synth.pangenome_text(60_000, 8 * world, 5).numpy(), 10, 100))
cases.append(("pangenome6", pkg.synth.pangenome_text(60_000, 8 * world, 5).numpy(), 6, 50))
cases.append(("random", pkg.synth.random_dna(300_000 * world, 6).numpy(), 10, 100))
a = pkg.synth.random_dna(40_000, 8).numpy()
cases.append(("nrun", np.concatenate([a, np.full(30_000 * world, ord("N"), np.uint8), a[:25_000]]), 10, 100))
cases.append(("pangenome16", pkg.synth.pangenome_text(60_000, 8 * world, 5).numpy(), 16, 500))
for name, text, w, p in cases:
    n = text.size
    cuts = [n * k // world + (7 * k if k < world else 0) for k in range(world + 1)]; cuts[-1] = n
    for mode in ("partition", "replicate"):
        try:
            got = T.emulate(pkg, text.tobytes(), cuts, w, p, mode, halo=4096, front=1 << 20)
            want = orc.parse(text.tobytes(), w, p)
            print(name, mode, {e: got[e] == getattr(want, e) for e in ("dict", "occ", "parse", "last", "sai")}, flush=True)
        except Exception as ex:
            print(name, mode, "EXC", repr(ex)[:300], flush=True)
