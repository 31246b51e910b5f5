"""Sharded prefix-free parsing: one process per GPU, the text split into contiguous shards.

World size 1 is the plain single-GPU parser.  (Multi-GPU: see ShardedParser docstring.)
"""
from __future__ import annotations

import numpy as np

from . import pfp


class ShardedParser:
    """Holds this rank's shard of the text in HBM and parses the whole text cooperatively.

    Reference analogue: the per-thread input ranges of pscan.hpp:114-165 / newscan.hpp:230-337.
    """

    def __init__(self, device: int, world: int = 1, rank: int = 0):
        self.device, self.world, self.rank = device, world, rank
        self.scanner = pfp.Scanner(device)
        self.text = None
        self.n_local = 0
        self.n_global = 0
        self.out = None

    def set_text(self, text):
        """text: CUDA uint8 tensor with this rank's contiguous shard of the global text."""
        self.text = text
        self.n_local = int(text.numel())
        if self.world == 1:
            self.n_global = self.n_local
        else:
            import torch
            import torch.distributed as dist
            sizes = torch.zeros(self.world, dtype=torch.int64, device=text.device)
            sizes[self.rank] = self.n_local
            dist.all_reduce(sizes)
            self.sizes = [int(x) for x in sizes.tolist()]
            self.n_global = sum(self.sizes)
            self.pos0 = sum(self.sizes[:self.rank])

    def release_text(self):
        self.text = None

    def parse_device(self, w=10, p=100, sai=True) -> dict:
        if self.world != 1:
            raise NotImplementedError("multi-GPU parse not wired yet")
        self.out = self.scanner.parse_device(self.text, w, p, sai=sai)
        return self.scanner.stats.as_dict()

    def parse_host(self, host_text, w=10, p=100, sai=True) -> dict:
        """host_text: pinned CPU uint8 tensor (this rank's shard)."""
        if self.world != 1:
            raise NotImplementedError("multi-GPU parse not wired yet")
        out = self.scanner.parse_host_ptr(host_text.data_ptr(), host_text.numel(), w, p, sai=sai)
        st = self.scanner.stats.as_dict()
        P, d = out.n_phrases, out.n_distinct
        st["h2d_bytes"] = int(st["n_text"])
        st["d2h_bytes"] = int(out.dict_bytes + 4 * d + 4 * P + P + (5 * P if sai else 0))
        self.host_out = out
        return st
