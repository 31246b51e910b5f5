import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as g
pkg = g.load_package()
text = pkg.synth.pangenome_text(40_000_000, 100, 2, device="cuda")
torch.cuda.synchronize()
for mode in ("", "rolling"):
    if mode: os.environ["PFPB200_K1"] = mode
    else: os.environ.pop("PFPB200_K1", None)
    sc = pkg.pfp.Scanner(0)
    for (w, p) in [(16, 100), (16, 500), (12, 100), (10, 100)]:
        for _ in range(3):
            out = sc.parse_device(text, w, p, sai=True)
        st = sc.stats.as_dict()
        print(json.dumps({"mode": mode or "interval", "w": w, "p": p, "ms_scan": round(st["ms_scan"], 3), "ms_total": round(st["ms_total"], 3), "phrases": out.n_phrases}))
    sc.close()
