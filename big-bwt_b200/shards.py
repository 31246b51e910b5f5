"""Sharded prefix-free parsing: one process per GPU, the text split into contiguous shards.

Reference analogue: the per-thread input ranges of pscan.hpp:114-165 / newscan.hpp:230-337 and
the shared dictionary of pscan.cpp:137-205.  Protocol of one parse (rank g, G ranks):

  1. halo     : rank g>0 receives the last HALO bytes of shard g-1 (>= w: w-1 for the window
                that ends at its first position, 1 for that phrase's `.last` byte)
  2. scan     : K1 on [halo|shard]; the rank owns the triggers whose END lies in its shard, with
                the sequential first-window rule (e >= w-1), so the union over ranks is exactly
                the sequential scanner's trigger set
  3. seams    : all-gather of (trigger count, last trigger).  The first phrase ending in shard g
                starts at (last trigger of any lower rank) - w + 1; if that lies before the halo,
                the missing head bytes are fetched peer to peer from the ranks that own them --
                however far back that is (a multi-megabyte run of N is ONE phrase that may cover
                several whole shards): the rank's buffer grows in front as needed
  4. words    : K2 + K3 on the shard -> local dictionary (fingerprint, length, count, pool bytes),
                `.last`/`.sai` of the shard
  5. merge    : mode "replicate": all-gather of the local dictionaries, every rank runs the global
                dedup + ranking (identical `.dict/.occ` everywhere).
                mode "partition": words are routed to the rank owning their lexicographic range
                (splitters from an all-gathered sample of first keys; identical words share a
                range, so one all-to-all serves dedup AND ranking); each owner dedups and ranks
                its range; global rank = local rank + number of distinct words in lower ranges;
                ranks travel back with a second all-to-all.  `.dict/.occ` are the per-rank pieces
                in rank order.
  6. remap    : `.parse` of the shard from the global rank of each local word (K5)

World size 1 short-circuits to the single-GPU parser.  The collectives are torch.distributed
(NCCL on GPUs; the same code runs over gloo with the CPU mock backend of tests/).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import pfp

HALO = 4096
FRONT = 1 << 20
MAX_RANKS = 64          # PFPB200_MAX_RANKS


# ------------------------------------------------------------------------------------------------
# pure planning helpers (unit-tested on CPU)
# ------------------------------------------------------------------------------------------------
def first_phrase_start(rank: int, n_trig: list[int], last_trig: list[int], w: int) -> int:
    """Global start of the first phrase that ends in shard `rank`: previous trigger - w + 1, or -1
    (the virtual 0x02 in front of the text) when no lower rank has a trigger."""
    for q in range(rank - 1, -1, -1):
        if n_trig[q] > 0:
            return last_trig[q] - w + 1
    return -1


def head_requests(pos0: list[int], n_local: list[int], n_trig: list[int], last_trig: list[int],
                  w: int, halo: int) -> list[tuple[int, int]]:
    """For every rank the global byte range [a, b) it still needs in front of its halo (a == b:
    nothing).  A rank without phrases (no trigger and not last) needs nothing."""
    G = len(pos0)
    out = []
    for g in range(G):
        has_phrase = n_trig[g] > 0 or g == G - 1
        have_from = max(0, pos0[g] - halo)
        a = b = have_from
        if g > 0 and has_phrase:
            fs = max(0, first_phrase_start(g, n_trig, last_trig, w))
            if fs < have_from:
                a = fs
        out.append((a, b))
    return out


def transfers(requests: list[tuple[int, int]], pos0: list[int], n_local: list[int]):
    """(src, dst, global_lo, global_hi) copies that satisfy the requests from the owning shards."""
    ops = []
    for dst, (a, b) in enumerate(requests):
        if a >= b:
            continue
        for src in range(len(pos0)):
            lo, hi = max(a, pos0[src]), min(b, pos0[src] + n_local[src])
            if lo < hi:
                ops.append((src, dst, lo, hi))
    return ops


class _DevView:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr, nbytes, typestr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


_TYPES = {torch.uint8: ("|u1", 1), torch.int32: ("<i4", 4), torch.int64: ("<i8", 8)}


def dev_tensor(ptr, count, dtype, device):
    if not ptr or count == 0:
        return torch.empty(0, dtype=dtype, device=device)
    ts, sz = _TYPES[dtype]
    return torch.as_tensor(_DevView(ptr, count * sz, ts, count), device=device)


# ------------------------------------------------------------------------------------------------
# GPU backend: the stage-level C ABI on torch tensors
# ------------------------------------------------------------------------------------------------
class CudaBackend:
    def __init__(self, device: int):
        self.scanner = pfp.Scanner(device)
        self.L, self.h = self.scanner.L, self.scanner.h
        self.dev = torch.device("cuda", device)
        self._declare()
        self.ms = {}
        self._stream = None
        self.use_current_stream()

    def use_current_stream(self):
        """Run the library on torch's current stream so that its kernels are ordered with the
        collectives and tensor ops around them."""
        s = torch.cuda.current_stream(self.dev).cuda_stream
        if s == 0:
            s = 1          # cudaStreamLegacy: the handle 0 would mean "the library's own stream"
        if s != self._stream:
            self.scanner.set_stream(s)
            self._stream = s

    def _declare(self):
        L = self.L
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.pfpb200_shard_scan.argtypes = [vp, C.POINTER(Shard), C.POINTER(pfp.Opts), C.POINTER(u64),
                                         C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float)]
        L.pfpb200_shard_words.argtypes = [vp, C.c_int64, C.POINTER(Words), C.POINTER(C.c_float)]
        L.pfpb200_dict_merge.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, u64, u32, u32,
                                         C.POINTER(Merged), C.POINTER(C.c_float)]
        L.pfpb200_shard_remap.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_float)]
        L.pfpb200_shard_first_keys.argtypes = [vp, C.POINTER(vp)]
        L.pfpb200_shard_route.argtypes = [vp, vp, u32, C.POINTER(Routed), C.POINTER(C.c_float)]
        L.pfpb200_dict_merge_words.argtypes = [vp, u64, vp, vp, u64, u32, u32, C.POINTER(Merged),
                                               C.POINTER(C.c_float)]
        L.pfpb200_shard_route_plan.argtypes = [vp, vp, u32, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp),
                                               C.POINTER(C.c_float)]
        L.pfpb200_shard_route_plan.restype = C.c_int
        L.pfpb200_shard_route_push.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float)]
        L.pfpb200_shard_route_push.restype = C.c_int
        L.pfpb200_dict_merge_words.restype = C.c_int
        L.pfpb200_dict_merge_begin.argtypes = [vp, u64, vp, C.POINTER(C.c_float)]
        L.pfpb200_dict_merge_begin.restype = C.c_int
        L.pfpb200_dict_merge_finish.argtypes = [vp, vp, u64, u32, u32, C.POINTER(Merged), C.POINTER(C.c_float)]
        L.pfpb200_dict_merge_finish.restype = C.c_int
        for f in (L.pfpb200_shard_scan, L.pfpb200_shard_words, L.pfpb200_dict_merge, L.pfpb200_shard_remap,
                  L.pfpb200_shard_first_keys, L.pfpb200_shard_route):
            f.restype = C.c_int

    def shard_scan(self, buf, buf_pos0, own_lo, own_hi, n_global, is_last, w, p, sai, verify=False):
        sh = Shard(buf.data_ptr(), buf.numel(), buf_pos0, own_lo, own_hi, n_global, 1 if is_last else 0, 0)
        o = pfp.Opts(w, p, (pfp.F_SAI if sai else 0) | (pfp.F_VERIFY if verify else 0), 0)
        n, first, last, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_float()
        self.scanner._check(self.L.pfpb200_shard_scan(self.h, C.byref(sh), C.byref(o), C.byref(n),
                                                      C.byref(first), C.byref(last), C.byref(ms)))
        self.ms["scan"] = ms.value
        return n.value, first.value, last.value

    def shard_words(self, first_start):
        wd, ms = Words(), C.c_float()
        self.scanner._check(self.L.pfpb200_shard_words(self.h, first_start, C.byref(wd), C.byref(ms)))
        self.ms["words"] = ms.value
        d, dev = wd.n_words, self.dev
        return {"n_words": d, "n_phrases": wd.n_phrases,
                "fpa": dev_tensor(wd.fpa, d, torch.int64, dev), "fpb": dev_tensor(wd.fpb, d, torch.int64, dev),
                "len": dev_tensor(wd.len, d, torch.int32, dev), "count": dev_tensor(wd.count, d, torch.int32, dev),
                "uwords": dev_tensor(wd.uwords, d, torch.int32, dev),
                "pool": dev_tensor(wd.pool, wd.pool_words, torch.int64, dev),
                "last": dev_tensor(wd.last, wd.n_phrases, torch.uint8, dev),
                "sai": dev_tensor(wd.sai, 5 * wd.n_phrases if wd.sai else 0, torch.uint8, dev)}

    def dict_merge(self, fpa, fpb, ln, count, uwords, pool, w, compress=False, verify=False):
        m, ms = Merged(), C.c_float()
        self.scanner._check(self.L.pfpb200_dict_merge(
            self.h, fpa.numel(), fpa.data_ptr(), fpb.data_ptr(), ln.data_ptr(), count.data_ptr(),
            uwords.data_ptr(), pool.data_ptr(), pool.numel(), w,
            (pfp.F_COMPRESS if compress else 0) | (pfp.F_VERIFY if verify else 0),
            C.byref(m), C.byref(ms)))
        self.ms["merge"] = ms.value
        dev = self.dev
        return {"n_distinct": m.n_distinct, "sum_word_len": m.sum_word_len,
                "dict": dev_tensor(m.dict, m.dict_bytes, torch.uint8, dev),
                "occ": dev_tensor(m.occ, m.n_distinct, torch.int32, dev),
                "rank_of_entry": dev_tensor(m.rank_of_entry, fpa.numel(), torch.int32, dev)}

    def first_keys(self, wd):
        out = C.c_void_p()
        self.scanner._check(self.L.pfpb200_shard_first_keys(self.h, C.byref(out)))
        return dev_tensor(out.value, wd["n_words"], torch.int64, self.dev)

    def sample_keys(self, max_samples: int) -> np.ndarray:
        """First keys of (up to) max_samples evenly spaced words of the local dictionary (host)."""
        self.L.pfpb200_shard_sample_keys.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        self.L.pfpb200_shard_sample_keys.restype = C.c_int
        out = np.zeros(max_samples, dtype=np.uint64)
        n = C.c_uint32()
        self.scanner._check(self.L.pfpb200_shard_sample_keys(self.h, max_samples, C.c_void_p(out.ctypes.data), C.byref(n)))
        return out[:n.value]

    def ranks_back(self, n_ranks, back, rank_base, n_words):
        """Global rank of every local word from the ranks that came back in routed order."""
        self.L.pfpb200_shard_ranks_back.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64),
                                                    C.POINTER(C.c_void_p)]
        self.L.pfpb200_shard_ranks_back.restype = C.c_int
        base = (C.c_uint64 * n_ranks)(*rank_base)
        out = C.c_void_p()
        self.scanner._check(self.L.pfpb200_shard_ranks_back(self.h, n_ranks, C.c_void_p(back.data_ptr()), base, C.byref(out)))
        return dev_tensor(out.value, n_words, torch.int32, self.dev)

    def route(self, wd, splitters: np.ndarray, n_ranks: int):
        rt, ms = Routed(), C.c_float()
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        self.scanner._check(self.L.pfpb200_shard_route(self.h, C.c_void_p(sp.ctypes.data if sp.size else 0),
                                                       n_ranks, C.byref(rt), C.byref(ms)))
        self.ms["route"] = ms.value
        d, dev = wd["n_words"], self.dev
        return {"words": dev_tensor(rt.words, 32 * d, torch.uint8, dev),     # pfpb200_word records
                "pool": dev_tensor(rt.pool, wd["pool"].numel(), torch.int64, dev),
                "perm": dev_tensor(rt.perm, d, torch.int32, dev),
                "words_to": [int(rt.words_to[q]) for q in range(n_ranks)],
                "pool_to": [int(rt.pool_to[q]) for q in range(n_ranks)]}

    def route_plan(self, wd, splitters: np.ndarray, n_ranks: int):
        """Owner of every word: counts per owner and the routed order; nothing is copied yet."""
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        wt, pt = (C.c_uint64 * MAX_RANKS)(), (C.c_uint64 * MAX_RANKS)()
        perm, ms = C.c_void_p(), C.c_float()
        self.scanner._check(self.L.pfpb200_shard_route_plan(self.h, C.c_void_p(sp.ctypes.data if sp.size else 0),
                                                            n_ranks, wt, pt, C.byref(perm), C.byref(ms)))
        self.ms["route"] = ms.value
        return {"perm": dev_tensor(perm.value, wd["n_words"], torch.int32, self.dev),
                "words_to": [int(wt[q]) for q in range(n_ranks)], "pool_to": [int(pt[q]) for q in range(n_ranks)]}

    def route_push(self, n_ranks, word_dst, pool_dst):
        """Store this rank's records / pool words straight into the owners' buffers (peer addresses)."""
        wd_, pd_ = (C.c_uint64 * MAX_RANKS)(*word_dst), (C.c_uint64 * MAX_RANKS)(*pool_dst)
        ms = C.c_float()
        self.scanner._check(self.L.pfpb200_shard_route_push(self.h, n_ranks, wd_, pd_, C.byref(ms)))
        self.ms["route"] = self.ms.get("route", 0.0) + ms.value

    def dict_merge_words(self, words, pool, w, compress=False, verify=False):
        m, ms = Merged(), C.c_float()
        n_in = words.numel() // 32
        self.scanner._check(self.L.pfpb200_dict_merge_words(
            self.h, n_in, words.data_ptr(), pool.data_ptr(), pool.numel(), w,
            (pfp.F_COMPRESS if compress else 0) | (pfp.F_VERIFY if verify else 0), C.byref(m), C.byref(ms)))
        self.ms["merge"] = ms.value
        dev = self.dev
        return {"n_distinct": m.n_distinct, "sum_word_len": m.sum_word_len,
                "dict": dev_tensor(m.dict, m.dict_bytes, torch.uint8, dev),
                "occ": dev_tensor(m.occ, m.n_distinct, torch.int32, dev),
                "rank_of_entry": dev_tensor(m.rank_of_entry, n_in, torch.int32, dev)}

    def dict_merge_begin(self, words):
        """Global dedup of the received 32-byte word records (no bytes needed yet)."""
        ms = C.c_float()
        self._merge_n_in = words.numel() // 32
        self.scanner._check(self.L.pfpb200_dict_merge_begin(self.h, self._merge_n_in, words.data_ptr(), C.byref(ms)))
        self.ms["merge"] = ms.value

    def dict_merge_finish(self, pool, w, compress=False, verify=False):
        """Ranking + .dict/.occ of the deduplicated words, once their bytes have arrived."""
        m, ms = Merged(), C.c_float()
        self.scanner._check(self.L.pfpb200_dict_merge_finish(
            self.h, pool.data_ptr(), pool.numel(), w,
            (pfp.F_COMPRESS if compress else 0) | (pfp.F_VERIFY if verify else 0), C.byref(m), C.byref(ms)))
        self.ms["merge"] = self.ms.get("merge", 0.0) + ms.value
        dev = self.dev
        return {"n_distinct": m.n_distinct, "sum_word_len": m.sum_word_len,
                "dict": dev_tensor(m.dict, m.dict_bytes, torch.uint8, dev),
                "occ": dev_tensor(m.occ, m.n_distinct, torch.int32, dev),
                "rank_of_entry": dev_tensor(m.rank_of_entry, self._merge_n_in, torch.int32, dev)}

    def shard_remap(self, rank_of_word, n_phrases):
        out, ms = C.c_void_p(), C.c_float()
        self.scanner._check(self.L.pfpb200_shard_remap(self.h, rank_of_word.data_ptr(), C.byref(out), C.byref(ms)))
        self.ms["remap"] = ms.value
        return dev_tensor(out.value, n_phrases, torch.int32, self.dev)


class Shard(C.Structure):
    _fields_ = [("d_buf", C.c_void_p), ("n_buf", C.c_uint64), ("buf_pos0", C.c_uint64),
                ("own_lo", C.c_uint64), ("own_hi", C.c_uint64), ("n_global", C.c_uint64),
                ("is_last", C.c_uint32), ("reserved", C.c_uint32)]


class Words(C.Structure):
    _fields_ = [("n_words", C.c_uint64), ("n_phrases", C.c_uint64), ("pool_words", C.c_uint64),
                ("fpa", C.c_void_p), ("fpb", C.c_void_p), ("len", C.c_void_p), ("count", C.c_void_p),
                ("uwords", C.c_void_p), ("pool", C.c_void_p), ("last", C.c_void_p), ("sai", C.c_void_p)]


class Routed(C.Structure):
    _fields_ = [("words", C.c_void_p), ("pool", C.c_void_p), ("perm", C.c_void_p),
                ("words_to", C.c_uint64 * 64), ("pool_to", C.c_uint64 * 64)]


class Merged(C.Structure):
    _fields_ = [("n_distinct", C.c_uint64), ("dict_bytes", C.c_uint64), ("sum_word_len", C.c_uint64),
                ("dict", C.c_void_p), ("occ", C.c_void_p), ("rank_of_entry", C.c_void_p)]


# ------------------------------------------------------------------------------------------------
# the sharded parser
# ------------------------------------------------------------------------------------------------
class PeerExchange:
    """All-to-all over NVLink peer memory: every rank stores its outgoing segments straight into
    the receivers' buffers (torch symmetric memory: the peers' allocations are mapped into this
    process), one DMA per peer on its own stream, bracketed by two device-side barriers.  Measured
    on 2 x B200, 0.42 GB per rank: 0.58 ms against 0.85 ms for NCCL's grouped send/recv; no
    staging buffers, no protocol, and the copies of different peers overlap."""

    def __init__(self, device, world, rank, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem, self.dev, self.world, self.rank, self.group = symm_mem, device, world, rank, group
        self.bufs = {}           # name -> (tensor, handle, capacity in elements)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(world)]

    def ensure(self, name, dtype, capacity):
        """Collective: every rank calls it with the same capacity."""
        cur = self.bufs.get(name)
        if cur is not None and cur[2] >= capacity:
            return
        cap = int(capacity * 1.25) + 4096
        t = self.symm_mem.empty(cap, dtype=dtype, device=self.dev)
        h = self.symm_mem.rendezvous(t, self.group)
        self.bufs[name] = (t, h, cap)

    def barrier(self, name):
        self.bufs[name][1].barrier()

    def local(self, name, count):
        return self.bufs[name][0][:count]

    def ptrs(self, name):
        """Device addresses of every rank's buffer `name` as mapped into this process."""
        return [int(p) for p in self.bufs[name][1].buffer_ptrs]

    def join(self):
        """The current stream waits for the copies issued by scatter(..., wait=False)."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)

    def scatter(self, name, send, send_off, send_cnt, dst_off, wait=True):
        """send[send_off[q] : +send_cnt[q]] -> rank q's buffer `name` at element dst_off[q].
        wait=False: the copies run on the side streams while the current stream goes on; join() later."""
        t, h, cap = self.bufs[name]
        cur = torch.cuda.current_stream(self.dev)
        for i in range(self.world):
            q = (self.rank + i) % self.world
            c = int(send_cnt[q])
            if c == 0:
                continue
            dst = h.get_buffer(q, (cap,), t.dtype)[int(dst_off[q]):int(dst_off[q]) + c]
            st = self.streams[i]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                dst.copy_(send[int(send_off[q]):int(send_off[q]) + c], non_blocking=True)
        if wait:
            for st in self.streams:
                cur.wait_stream(st)


class ShardedParser:
    """Holds this rank's shard of the text and parses the whole text with the other ranks."""

    def __init__(self, device: int | None, world: int = 1, rank: int = 0, backend=None, mode="replicate",
                 halo: int = HALO, front: int = FRONT):
        self.world, self.rank, self.mode = world, rank, mode
        self.halo, self.front_cap = halo, front
        self.backend = backend if backend is not None else CudaBackend(device)
        self.scanner = getattr(self.backend, "scanner", None)
        self.buf = None
        self.n_local = self.n_global = self.pos0 = 0
        self.result = None
        self._peer = None          # PeerExchange once it has been set up (GPU ranks only)
        self._peer_failed = False

    # -- collectives --------------------------------------------------------------------------------
    def _all_gather_i64(self, vals):
        import torch.distributed as dist
        t = torch.tensor(vals, dtype=torch.int64, device=self.buf.device)
        out = torch.empty(self.world * len(vals), dtype=torch.int64, device=self.buf.device)
        dist.all_gather_into_tensor(out, t)
        return out.view(self.world, len(vals)).tolist()

    def _all_gather_v(self, t, counts):
        """Concatenation over ranks of 1-D tensors of different lengths (padded all-gather)."""
        import torch.distributed as dist
        mx = max(counts) if counts else 0
        if mx == 0:
            return torch.empty(0, dtype=t.dtype, device=t.device)
        send = t
        if t.numel() != mx:
            send = torch.empty(mx, dtype=t.dtype, device=t.device)
            send[:t.numel()] = t
        out = torch.empty(self.world * mx, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, send.contiguous())
        if all(c == mx for c in counts):
            return out
        return torch.cat([out[q * mx:q * mx + c] for q, c in enumerate(counts)])

    def _p2p(self, ops):
        """ops: (src, dst, global_lo, global_hi) byte ranges; executes this rank's part."""
        import torch.distributed as dist
        todo = []
        for src, dst, lo, hi in ops:
            if src == dst:
                continue
            if self.rank == src:
                a = self.front + (lo - self.pos0)
                todo.append(dist.P2POp(dist.isend, self.buf[a:a + (hi - lo)], dst))
            elif self.rank == dst:
                a = self.front - (self.pos0 - lo)          # >= 0: the buffer was grown to hold the head
                todo.append(dist.P2POp(dist.irecv, self.buf[a:a + (hi - lo)], src))
        if todo:
            for r in dist.batch_isend_irecv(todo):
                r.wait()

    # -- text ------------------------------------------------------------------------------------------
    def set_text(self, text):
        """text: this rank's contiguous shard (uint8 tensor on the compute device)."""
        self.n_local = int(text.numel())
        dev = text.device
        if self.world == 1:
            self.buf, self.front = text, 0
            self.n_global, self.pos0 = self.n_local, 0
            self.sizes = [self.n_local]
            return
        import torch.distributed as dist
        sizes = torch.zeros(self.world, dtype=torch.int64, device=dev)
        sizes[self.rank] = self.n_local
        dist.all_reduce(sizes)
        self.sizes = [int(x) for x in sizes.tolist()]
        self.n_global = sum(self.sizes)
        self.starts = [sum(self.sizes[:q]) for q in range(self.world)]
        self.pos0 = self.starts[self.rank]
        self.front = min(self.front_cap, self.pos0)
        self.buf = torch.empty(self.front + self.n_local, dtype=torch.uint8, device=dev)
        self.buf[self.front:].copy_(text)

    def release_text(self):
        self.buf = None

    # -- parse --------------------------------------------------------------------------------------------
    def parse_device(self, w=10, p=100, sai=True, compress=False, verify=False) -> dict:
        if self.world == 1 and self.scanner is not None:
            self.backend.use_current_stream()
            self.out = self.scanner.parse_device(self.buf, w, p, sai=sai, compress=compress, verify=verify)
            return self.scanner.stats.as_dict()
        return self._parse_sharded(w, p, sai, compress, verify)

    def parse_host(self, host_text, w=10, p=100, sai=True) -> dict:
        """host_text: pinned CPU uint8 tensor holding this rank's shard."""
        if self.world == 1:
            self.backend.use_current_stream()
            out = self.scanner.parse_host_ptr(host_text.data_ptr(), host_text.numel(), w, p, sai=sai)
            st = self.scanner.stats.as_dict()
            P, d = out.n_phrases, out.n_distinct
            st["h2d_bytes"] = int(st["n_text"])
            st["d2h_bytes"] = int(out.dict_bytes + 4 * d + 4 * P + P + (5 * P if sai else 0))
            self.host_out = out
            return st
        # sharded: H2D of the shard, cooperative parse, D2H of this rank's output pieces
        dev = self.backend.dev
        if self.buf is None or self.buf.numel() != self.front + host_text.numel():
            self.set_text(torch.empty(host_text.numel(), dtype=torch.uint8, device=dev))
        self.buf[self.front:].copy_(host_text, non_blocking=True)
        st = self._parse_sharded(w, p, sai, False)
        r = self.result
        outs = {"parse": r["parse"], "last": r["last"], "sai": r["sai"]}
        if self.mode == "partition" or self.rank == 0:
            outs.update({"dict": r["dict"], "occ": r["occ"]})
        # device -> pinned host buffers kept across calls (grown on demand), one sync at the end
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        self.host_pieces = {}
        nbytes = 0
        for k, t in outs.items():
            t = t.contiguous().view(torch.uint8) if t.dtype != torch.uint8 else t.contiguous()
            buf = self._pinned.get(k)
            if buf is None or buf.numel() < t.numel():
                buf = torch.empty(int(t.numel() * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
                self._pinned[k] = buf
            buf[:t.numel()].copy_(t, non_blocking=True)
            self.host_pieces[k] = buf[:t.numel()]
            nbytes += t.numel()
        torch.cuda.current_stream(dev).synchronize()
        st["h2d_bytes"] = int(host_text.numel())
        st["d2h_bytes"] = int(nbytes)
        return st

    def _mark(self, name):
        """Phase boundary for the timeline in the returned stats (CUDA events, no synchronisation)."""
        if self.buf is not None and self.buf.is_cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.buf.device))
            self._marks.append((name, ev))

    def _grow_front(self, need):
        """More room in front of the shard (the head of a phrase that started `need` bytes before
        it): new buffer, old content (halo included) at its end.  Local: no collective."""
        need = (need + 4095) // 4096 * 4096
        need = min(need, self.pos0)
        new = torch.empty(need + self.n_local, dtype=torch.uint8, device=self.buf.device)
        new[need - self.front:].copy_(self.buf)
        self.buf, self.front = new, need

    def _parse_sharded(self, w, p, sai, compress, verify=False) -> dict:
        be, G, g = self.backend, self.world, self.rank
        if hasattr(be, "use_current_stream"):
            be.use_current_stream()
        self._marks = []
        self._mark("start")
        if w > self.halo:
            raise ValueError("window larger than the shard halo")
        if min(self.sizes[:-1]) < self.halo:
            raise ValueError("shards must be at least HALO bytes long")
        # 1. halo
        halo_reqs = [(max(0, self.starts[q] - self.halo), self.starts[q]) for q in range(G)]
        self._p2p(transfers(halo_reqs, self.starts, self.sizes))
        self._mark("halo")
        # 2. scan
        def scan():
            return be.shard_scan(self.buf, self.pos0 - self.front, self.pos0, self.pos0 + self.n_local,
                                 self.n_global, g == G - 1, w, p, sai, verify)
        n_trig, first, last = scan()
        self._mark("scan")
        # 3. seams (every rank derives the same plan from the all-gathered triggers, so a rank
        #    never waits for a transfer its peer did not post)
        info = self._all_gather_i64([n_trig, last])
        nts, lasts = [r[0] for r in info], [r[1] for r in info]
        reqs = head_requests(self.starts, self.sizes, nts, lasts, w, self.halo)
        a, b = reqs[g]
        if a < b and self.pos0 - a > self.front:
            # the phrase straddling into this shard starts before the reserved front: grow the
            # buffer and scan it again (the library keeps pointers into the scanned buffer)
            self._grow_front(self.pos0 - a)
            n_trig, first, last = scan()
        self._p2p(transfers(reqs, self.starts, self.sizes))
        fs = first_phrase_start(g, nts, lasts, w)
        self._mark("seams")
        # 4. words
        wd = be.shard_words(fs)
        self._mark("words")
        # 5. merge
        if self.mode == "replicate":
            res = self._merge_replicated(wd, w, compress, verify)
        else:
            res = self._merge_partitioned(wd, w, compress, verify)
        # 6. remap
        parse = be.shard_remap(res["rank_of_word"], wd["n_phrases"])
        self._mark("remap")
        self.result = {"dict": res["dict"], "occ": res["occ"], "parse": parse, "last": wd["last"],
                       "sai": wd["sai"], "n_distinct": res["n_distinct"], "dict_offset": res.get("dict_offset", 0)}
        t = res["totals"]
        P, d, db, sl = t["n_phrases"], t["n_distinct"], t["dict_bytes"], t["sum_word_len"]
        ms = dict(getattr(be, "ms", {}))
        st = {"n_text": self.n_global, "n_phrases": P, "n_distinct": d, "dict_bytes": db, "sum_word_len": sl,
              "alg_bytes": self.n_global + 4 * P + P + (5 * P if sai else 0) + db + 4 * d,
              "rank_rounds": 0,
              "launches": int(be.L.pfpb200_launch_count(be.h)) if hasattr(be, "L") else 0,
              "ms_scan": ms.get("scan", 0.0), "ms_hash": ms.get("words", 0.0), "ms_rank": ms.get("merge", 0.0),
              "ms_remap": ms.get("remap", 0.0)}
        if self._marks:
            self._marks[-1][1].synchronize()
            for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
                st["ms_phase_" + name] = st.get("ms_phase_" + name, 0.0) + a.elapsed_time(b)
        return st

    def _merge_replicated(self, wd, w, compress, verify=False):
        be, G, g = self.backend, self.world, self.rank
        sizes = self._all_gather_i64([wd["n_words"], wd["pool"].numel()])
        nw, npool = [r[0] for r in sizes], [r[1] for r in sizes]
        cat = {k: self._all_gather_v(wd[k], nw) for k in ("fpa", "fpb", "len", "count", "uwords")}
        pool = self._all_gather_v(wd["pool"], npool)
        m = be.dict_merge(cat["fpa"], cat["fpb"], cat["len"], cat["count"], cat["uwords"], pool, w, compress, verify)
        base = sum(nw[:g])
        P = sum(r[0] for r in self._all_gather_i64([wd["n_phrases"]]))
        return {"dict": m["dict"], "occ": m["occ"], "n_distinct": m["n_distinct"],
                "rank_of_word": m["rank_of_entry"][base:base + nw[g]].contiguous(),
                "totals": {"n_phrases": P, "n_distinct": m["n_distinct"],
                           "dict_bytes": int(m["dict"].numel()), "sum_word_len": m["sum_word_len"]}}

    def _all_to_all_v(self, send, send_counts, recv_counts):
        """Variable-size all-to-all of a 1-D tensor grouped by destination (grouped P2P)."""
        import torch.distributed as dist
        out = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
        so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
        ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
        todo = []
        for q in range(self.world):
            if q == self.rank:
                out[ro[q]:ro[q + 1]] = send[so[q]:so[q + 1]]
                continue
            if send_counts[q]:
                todo.append(dist.P2POp(dist.isend, send[so[q]:so[q + 1]], q))
            if recv_counts[q]:
                todo.append(dist.P2POp(dist.irecv, out[ro[q]:ro[q + 1]], q))
        if todo:
            for r in dist.batch_isend_irecv(todo):
                r.wait()
        return out

    def _peer_exchange(self, dev):
        """The NVLink peer-memory exchange, or None (CPU ranks, or symmetric memory unavailable:
        the NCCL grouped send/recv path is used instead)."""
        if self._peer is not None or self._peer_failed:
            return self._peer
        import os
        if dev.type != "cuda" or os.environ.get("PFPB200_EXCHANGE", "peer") == "nccl":
            self._peer_failed = True
            return None
        import torch.distributed as dist
        ok = 1
        try:
            self._peer = PeerExchange(dev, self.world, self.rank, dist.group.WORLD)
            self._peer.ensure("words", torch.uint8, 1 << 20)
        except Exception:  # noqa: BLE001
            self._peer, ok = None, 0
        flag = torch.tensor([ok], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)           # all ranks or none
        if int(flag.item()) == 0:
            self._peer, self._peer_failed = None, True
        return self._peer

    SAMPLE = 1024

    def _splitters(self, wd):
        """G-1 range splitters (big-endian first-8-byte keys) from an all-gathered sample."""
        be, G = self.backend, self.world
        d = wd["n_words"]
        dev = self.buf.device
        samp = torch.zeros(self.SAMPLE + 1, dtype=torch.int64, device=dev)
        k = 0
        if d and hasattr(be, "sample_keys"):  # keys of the sampled words only, straight to the host and back
            pick = torch.from_numpy(be.sample_keys(self.SAMPLE).view(np.int64))
            k = int(pick.numel())
            samp[:k] = pick.to(dev)
        elif d:
            keys = be.first_keys(wd)          # uid order = creation order: any stride is a sample
            step = max(1, d // self.SAMPLE)
            pick = keys[::step][:self.SAMPLE]
            k = int(pick.numel())
            samp[:k] = pick
        samp[self.SAMPLE] = k
        import torch.distributed as dist
        allv = torch.empty(G * (self.SAMPLE + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allv, samp)
        a = allv.cpu().numpy().view(np.uint64).reshape(G, self.SAMPLE + 1)
        ks = [int(a[q, self.SAMPLE]) for q in range(G)]
        vals = np.sort(np.concatenate([a[q, :ks[q]] for q in range(G)])) if sum(ks) else np.zeros(0, np.uint64)
        if vals.size == 0:
            return np.zeros(G - 1, dtype=np.uint64)
        return np.array([vals[min(vals.size - 1, (q + 1) * vals.size // G)] for q in range(G - 1)], dtype=np.uint64)

    def _merge_partitioned(self, wd, w, compress, verify=False):
        be, G, g = self.backend, self.world, self.rank
        dev = self.buf.device
        sp = self._splitters(wd)
        self._mark("splitters")
        peer = self._peer_exchange(dev)
        import os
        fused = peer is not None and os.environ.get("PFPB200_EXCHANGE", "peer") == "push"
        if fused:
            # routing fused with the exchange: the route kernel stores this rank's records and pool
            # words straight into the owners' buffers over NVLink (rank g's part lands behind
            # those of the lower ranks; the all-gathered count matrix M gives every offset)
            rt = be.route_plan(wd, sp, G) if wd["n_words"] else None
        else:
            rt = be.route(wd, sp, G) if wd["n_words"] else None
        words_to = rt["words_to"] if rt else [0] * G
        pool_to = rt["pool_to"] if rt else [0] * G
        M = self._all_gather_i64(list(words_to) + list(pool_to))           # M[src] = [words_to.., pool_to..]
        recv_w = [M[q][g] for q in range(G)]
        recv_p = [M[q][G + g] for q in range(G)]
        self._mark("route")
        split = False                                   # pool bytes overlapped with the dedup of the records
        if peer is not None:
            need_w = max(sum(M[src][q] for src in range(G)) for q in range(G))
            need_p = max(sum(M[src][G + q] for src in range(G)) for q in range(G))
            need_r = max(sum(M[src][:G]) for src in range(G))
            peer.ensure("words", torch.uint8, 32 * need_w)
            peer.ensure("pool", torch.int64, need_p)
            peer.ensure("ranks", torch.int32, need_r)
            w_off = [sum(M[src][q] for src in range(g)) for q in range(G)]        # my slot in owner q's buffer
            p_off = [sum(M[src][G + q] for src in range(g)) for q in range(G)]
            peer.barrier("words")                       # the owners are done with the previous parse
            if fused:
                if rt:
                    wbase, pbase = peer.ptrs("words"), peer.ptrs("pool")
                    be.route_push(G, [wbase[q] + 32 * w_off[q] for q in range(G)],
                                  [pbase[q] + 8 * p_off[q] for q in range(G)])
            else:
                # one DMA per owner and buffer, each on its own stream (measured faster than SM stores
                # over NVLink: 0.42 GB in 0.58 ms against 1.39 ms on 2 GPUs)
                so_w = np.concatenate([[0], np.cumsum(words_to)]).astype(np.int64)
                so_p = np.concatenate([[0], np.cumsum(pool_to)]).astype(np.int64)
                send_words = rt["words"] if rt else torch.empty(0, dtype=torch.uint8, device=dev)
                send_pool = rt["pool"] if rt else torch.empty(0, dtype=torch.int64, device=dev)
                peer.scatter("words", send_words, 32 * so_w, [32 * c for c in words_to], [32 * o for o in w_off])
                split = hasattr(be, "dict_merge_begin") and hasattr(peer, "join")
                # the pool bytes travel on the side streams while the owners dedup the word records
                peer.scatter("pool", send_pool, so_p, pool_to, p_off, **({"wait": False} if split else {}))
            peer.barrier("words")                       # the word records addressed to me have landed
            got_words = peer.local("words", 32 * sum(recv_w))
            got_pool = peer.local("pool", sum(recv_p))
            if not fused and split:
                self._mark("exchange")
                be.dict_merge_begin(got_words)
                peer.join()
                peer.barrier("words")                   # ... and now the pool bytes
        else:
            # NCCL grouped send/recv: the 32-byte word records and the pool bytes
            send_words = rt["words"] if rt else torch.empty(0, dtype=torch.uint8, device=dev)
            send_pool = rt["pool"] if rt else torch.empty(0, dtype=torch.int64, device=dev)
            got_words = self._all_to_all_v(send_words, [32 * c for c in words_to], [32 * c for c in recv_w])
            got_pool = self._all_to_all_v(send_pool, pool_to, recv_p)
        if peer is not None and not fused and split:
            m = be.dict_merge_finish(got_pool, w, compress, verify)
        else:
            self._mark("exchange")
            m = be.dict_merge_words(got_words, got_pool, w, compress, verify)
        self._mark("merge")
        piece = m["dict"] if g == G - 1 else m["dict"][:-1]                 # only the last piece ends in 0x00
        tot = self._all_gather_i64([m["n_distinct"], int(piece.numel()), m["sum_word_len"], wd["n_phrases"]])
        nd = [r[0] for r in tot]
        offset = sum(nd[:g])
        c_back = peer is not None and hasattr(be, "ranks_back")     # offsets and un-routing in one library kernel
        ranks = m["rank_of_entry"]
        if not c_back and ranks.numel():
            ranks = ranks + offset
        if peer is not None:
            # the ranks of the records that came from rank q go back to q, behind what q routed to
            # the owners below me
            ro = np.concatenate([[0], np.cumsum(recv_w)]).astype(np.int64)
            peer.scatter("ranks", ranks.to(torch.int32), ro, recv_w,
                         [sum(M[q][:g]) for q in range(G)])
            peer.barrier("words")
            back = peer.local("ranks", wd["n_words"])
        else:
            back = self._all_to_all_v(ranks.to(torch.int32), recv_w, words_to)  # routed order
        if c_back and rt and not fused:
            rank_of_word = be.ranks_back(G, back, [sum(nd[:q]) for q in range(G)], wd["n_words"])
        else:
            if c_back and back.numel():         # (fused push keeps its own routing state: offsets here)
                so = np.concatenate([[0], np.cumsum(words_to)]).astype(np.int64)
                for q in range(G):
                    back[so[q]:so[q + 1]] += sum(nd[:q])
            rank_of_word = torch.empty(wd["n_words"], dtype=torch.int32, device=dev)
            if rt:
                rank_of_word[rt["perm"].long()] = back
        self._mark("ranks_back")
        return {"dict": piece, "occ": m["occ"], "n_distinct": sum(nd), "rank_of_word": rank_of_word,
                "totals": {"n_phrases": sum(r[3] for r in tot), "n_distinct": sum(nd),
                           "dict_bytes": sum(r[1] for r in tot), "sum_word_len": sum(r[2] for r in tot)}}

    # -- results ----------------------------------------------------------------------------------------------
    def gather_files(self):
        """All five streams of the whole text as bytes on every rank (tests / small inputs)."""
        import torch.distributed as dist
        r = self.result

        def cat_over_ranks(t):
            t = t.contiguous().view(torch.uint8) if t.dtype != torch.uint8 else t.contiguous()
            n = self._all_gather_i64([t.numel()])
            return bytes(self._all_gather_v(t, [x[0] for x in n]).cpu().numpy().tobytes())
        out = {k: cat_over_ranks(r[k]) for k in ("parse", "last", "sai")}
        if self.mode == "partition":
            out["dict"] = cat_over_ranks(r["dict"])
            out["occ"] = cat_over_ranks(r["occ"])
        else:
            out["dict"] = bytes(r["dict"].cpu().numpy().tobytes())
            out["occ"] = bytes(r["occ"].contiguous().view(torch.uint8).cpu().numpy().tobytes())
        return out
