#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
          --master-port 29511 tools/multi_gpu_check.py [replicate|partition]
Every rank parses its shard with the CUDA backend; rank 0 compares the gathered five streams with
the single-process oracle on the whole text."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import pfp_oracle as orc  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "replicate"
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = load_package()
    job = pkg.shards.ShardedParser(local, world, rank, mode=mode)
    ok = True
    cases = [("pangenome", 10, 100), ("pangenome", 6, 50), ("random", 10, 100), ("nrun", 10, 100), ("pangenome", 16, 500)]
    for name, w, p in cases:
        if name == "pangenome":
            text = pkg.synth.pangenome_text(60_000, 8 * world, 5).numpy()
        elif name == "random":
            text = pkg.synth.random_dna(300_000 * world, 6).numpy()
        else:
            a = pkg.synth.random_dna(40_000, 8).numpy()
            text = np.concatenate([a, np.full(30_000 * world, ord("N"), np.uint8), a[:25_000]])
        n = text.size
        cuts = [n * k // world + (7 * k if k < world else 0) for k in range(world + 1)]
        cuts[-1] = n
        shard = torch.from_numpy(text[cuts[rank]:cuts[rank + 1]].copy()).to(dev)
        job.set_text(shard)
        st = job.parse_device(w, p, sai=True)
        files = job.gather_files()
        if rank == 0:
            want = orc.parse(text.tobytes(), w, p)
            res = {e: files[e] == getattr(want, e) for e in ("dict", "occ", "parse", "last", "sai")}
            good = all(res.values()) and st["n_phrases"] == want.n_phrases and st["n_distinct"] == want.n_distinct
            ok &= good
            print(f"[{mode} x{world}] {name} w={w} p={p} n={n}: {'OK' if good else 'MISMATCH ' + str(res)} "
                  f"phrases={st['n_phrases']} distinct={st['n_distinct']}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
