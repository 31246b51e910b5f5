import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
from oracle import pfp_oracle as orc
pkg = load_package()
sc = pkg.pfp.Scanner(0)
for n in (1000, 40000):
    text = pkg.synth.pangenome_text(n // 10, 10, 3).numpy().tobytes()
    got = sc.parse_host(text, 10, 100); want = orc.parse(text, 10, 100)
    gw = got.dict[:-1].split(b"\x01")[:-1]; ww = want.dict[:-1].split(b"\x01")[:-1]
    print(n, "dict equal", got.dict == want.dict, "len", len(gw), len(ww), "same set", set(gw) == set(ww), "sorted", gw == sorted(gw))
    bad = [i for i in range(min(len(gw), len(ww))) if gw[i] != ww[i]]
    print(" first bad", bad[:5])
    for i in bad[:3]:
        print("  got ", gw[i][:60], len(gw[i])); print("  want", ww[i][:60], len(ww[i]))
    go = np.frombuffer(got.occ, np.uint32); wo = np.frombuffer(want.occ, np.uint32)
    print(" occ sum", go.sum(), wo.sum(), "occ equal", np.array_equal(go, wo))
