import sys, os
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as g
pkg = g.load_package()
sc = pkg.pfp.Scanner(0)
text = pkg.synth.random_dna(8_000_000_000, 4, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    out = sc.parse_device(text, 10, 100, sai=True)
print({k: round(v, 2) for k, v in sc.stats.as_dict().items() if k.startswith("ms_")})
