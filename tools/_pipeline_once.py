import sys, os
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as g
pkg = g.load_package()
sc = pkg.pfp.Scanner(0)
text = pkg.synth.pangenome_text(40_000_000, 100, 2, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    r, out, bp = sc.bwt_of_text(text, 10, 100, flags=0)
print("pfbwt ms", r.ms_total, "bwtparse ms", bp.ms_total)
