"""CPU stand-in for the stage-level C ABI (big-bwt_b200/shards.py CudaBackend), built on numpy and
the oracle's window hash.  It lets the multi-rank protocol of ShardedParser (halo / head fetch,
seam resolution, dictionary merge, rank offsets) run over gloo without a GPU.  Test code only."""
import hashlib

import numpy as np
import torch

from oracle import pfp_oracle as orc


def _fp(word: bytes):
    h = hashlib.blake2b(word, digest_size=16).digest()
    return (int.from_bytes(h[:8], "little", signed=True), int.from_bytes(h[8:], "little", signed=True))


def _pack(words):
    """zero padded 8-byte pool + per-word 8-byte word counts"""
    uw = [(len(x) + 7) // 8 for x in words]
    raw = b"".join(x + b"\0" * (8 * k - len(x)) for x, k in zip(words, uw))
    return np.frombuffer(raw, dtype=np.int64).copy(), uw


def _unpack(pool, uwords, lens):
    raw = pool.tobytes()
    out, off = [], 0
    for k, ln in zip(uwords, lens):
        out.append(raw[off:off + ln])
        off += 8 * k
    return out


class MockBackend:
    def __init__(self):
        self.ms = {}

    def shard_scan(self, buf, buf_pos0, own_lo, own_hi, n_global, is_last, w, p, sai, verify=False):
        self.buf = buf.numpy()
        self.buf_pos0, self.own = buf_pos0, (own_lo, own_hi)
        self.n_global, self.is_last, self.w, self.sai = n_global, is_last, w, sai
        # windows need w-1 bytes in front of the first owned position
        start = max(buf_pos0, own_lo - (w - 1))
        seg = self.buf[start - buf_pos0: own_hi - buf_pos0].tobytes()
        e = orc.triggers(seg, w, p, flags=orc.THREADED_RULE).astype(np.int64) + start
        # the threaded rule drops only a trigger at the segment's very first window, which ends
        # at start + w - 1: re-check it by hand when it is owned and not the text's zero-padded one
        first = start + w - 1
        if first < own_hi and len(seg) >= w and orc.window_hash(seg[:w]) % p == 0:
            e = np.concatenate([[first], e])
        e = e[(e >= own_lo) & (e < own_hi) & (e >= w - 1)]
        self.trig = e
        if len(e) == 0:
            return 0, 0, 0
        return len(e), int(e[0]), int(e[-1])

    def _byte(self, g):
        if g < 0 or g >= self.n_global:
            return 2
        return int(self.buf[g - self.buf_pos0])

    def shard_words(self, first_start):
        w = self.w
        ends = list(self.trig) + ([self.n_global + w - 1] if self.is_last else [])
        words, index, uid, last, sai = [], {}, [], bytearray(), bytearray()
        prev = None
        for j, e in enumerate(ends):
            s0 = first_start if j == 0 else prev - w + 1
            ph = bytes(self._byte(x) for x in range(s0, e + 1))
            if ph not in index:
                index[ph] = len(words)
                words.append(ph)
            uid.append(index[ph])
            last.append(self._byte(e - w))
            sai += int(e + 1).to_bytes(5, "little")
            prev = e
        self.words = words
        self.uid = np.array(uid, dtype=np.int64)
        cnt = np.bincount(self.uid, minlength=len(words)) if words else np.zeros(0, np.int64)
        fps = [_fp(x) for x in words]
        pool, uw = _pack(words)
        t = torch.from_numpy
        return {"n_words": len(words), "n_phrases": len(ends),
                "fpa": t(np.array([f[0] for f in fps], dtype=np.int64)),
                "fpb": t(np.array([f[1] for f in fps], dtype=np.int64)),
                "len": t(np.array([len(x) for x in words], dtype=np.int32)),
                "count": t(cnt.astype(np.int32)), "uwords": t(np.array(uw, dtype=np.int32)),
                "pool": t(pool), "last": t(np.frombuffer(bytes(last), np.uint8).copy()),
                "sai": t(np.frombuffer(bytes(sai) if self.sai else b"", np.uint8).copy())}

    def first_keys(self, wd):
        ks = [int.from_bytes(x[:8].ljust(8, b"\0"), "big") for x in self.words]
        return torch.from_numpy(np.array(ks, dtype=np.uint64).view(np.int64))

    def route(self, wd, splitters, n_ranks):
        ks = self.first_keys(wd).numpy().view(np.uint64)
        dest = np.array([int(np.sum(splitters <= k)) for k in ks], dtype=np.int64)
        perm = np.argsort(dest, kind="stable")
        words = [self.words[i] for i in perm]
        pool, uw = _pack(words)
        rec = np.zeros(len(words), dtype=[("fpa", "<i8"), ("fpb", "<i8"), ("len", "<u4"), ("count", "<u4"),
                                          ("uwords", "<u4"), ("pad", "<u4")])
        for key in ("fpa", "fpb", "len", "count"):
            rec[key] = wd[key].numpy()[perm]
        rec["uwords"] = uw
        t = torch.from_numpy
        return {"words": t(rec.view(np.uint8).copy()), "pool": t(pool), "perm": t(perm.astype(np.int32)),
                "words_to": [int(np.sum(dest == q)) for q in range(n_ranks)],
                "pool_to": [int(sum(uw[i] for i in range(len(uw)) if dest[perm[i]] == q)) for q in range(n_ranks)]}

    def dict_merge_words(self, words, pool, w, compress=False, verify=False):
        rec = words.numpy().view([("fpa", "<i8"), ("fpb", "<i8"), ("len", "<u4"), ("count", "<u4"),
                                  ("uwords", "<u4"), ("pad", "<u4")])
        t = torch.from_numpy
        return self.dict_merge(t(rec["fpa"].copy()), t(rec["fpb"].copy()), t(rec["len"].astype(np.int32)),
                               t(rec["count"].astype(np.int32)), t(rec["uwords"].astype(np.int32)), pool, w, compress)

    def dict_merge(self, fpa, fpb, ln, count, uwords, pool, w, compress=False, verify=False):
        lens = ln.numpy().tolist()
        words = _unpack(pool.numpy(), uwords.numpy().tolist(), lens)
        keys = list(zip(fpa.numpy().tolist(), fpb.numpy().tolist(), lens))
        tot, rep = {}, {}
        for k, c, wd in zip(keys, count.numpy().tolist(), words):
            tot[k] = tot.get(k, 0) + c
            assert rep.setdefault(k, wd) == wd, "fingerprint collision in the mock"
        order = sorted(tot, key=lambda k: rep[k])
        rank = {k: i + 1 for i, k in enumerate(order)}
        d = b"".join(rep[k] + b"\x01" for k in order) + b"\x00"
        t = torch.from_numpy
        return {"n_distinct": len(order), "sum_word_len": sum(len(rep[k]) for k in order),
                "dict": t(np.frombuffer(d, np.uint8).copy()),
                "occ": t(np.array([tot[k] for k in order], dtype=np.int32)),
                "rank_of_entry": t(np.array([rank[k] for k in keys], dtype=np.int32))}

    def shard_remap(self, rank_of_word, n_phrases):
        r = rank_of_word.numpy()
        return torch.from_numpy(r[self.uid].astype(np.int32)) if n_phrases else torch.zeros(0, dtype=torch.int32)


class SharedMemExchange:
    """CPU stand-in for shards.PeerExchange (NVLink peer memory): every rank's receive buffers are
    torch tensors in POSIX shared memory, mapped into every process -- so the offset arithmetic
    of the peer-memory exchange (my slot in owner q's buffer, the ranks travelling back) runs in
    the gloo tests exactly as it does over symmetric memory on the GPUs."""

    def __init__(self, bufs, rank, world):
        self.bufs, self.rank, self.world = bufs, rank, world    # name -> [tensor of rank 0, 1, ...]

    @staticmethod
    def allocate(world, words=1 << 20, pool=1 << 18, ranks=1 << 16):
        mk = lambda n, dt: [torch.zeros(n, dtype=dt).share_memory_() for _ in range(world)]  # noqa: E731
        return {"words": mk(words, torch.uint8), "pool": mk(pool, torch.int64), "ranks": mk(ranks, torch.int32)}

    def ensure(self, name, dtype, capacity):
        t = self.bufs[name][self.rank]
        assert t.dtype == dtype and capacity <= t.numel(), (name, capacity, t.numel())

    def barrier(self, name):
        import torch.distributed as dist
        dist.barrier()

    def local(self, name, count):
        return self.bufs[name][self.rank][:count]

    def scatter(self, name, send, send_off, send_cnt, dst_off):
        for q in range(self.world):
            c = int(send_cnt[q])
            if c:
                self.bufs[name][q][int(dst_off[q]):int(dst_off[q]) + c] = send[int(send_off[q]):int(send_off[q]) + c]
