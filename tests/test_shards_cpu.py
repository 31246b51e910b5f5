"""The N>1 path on the CPU: ShardedParser's protocol over gloo (world sizes 2 and 3) with the
mock stage backend, checked byte-for-byte against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, FILES
from oracle import pfp_oracle as orc


def test_first_phrase_start_and_head_requests(pkg):
    sh = pkg.shards
    # ranks 0..3, rank 1 and 2 have no trigger
    nt, lt = [3, 0, 0, 5], [950, 0, 0, 3900]
    assert sh.first_phrase_start(0, nt, lt, 10) == -1
    assert sh.first_phrase_start(1, nt, lt, 10) == 941
    assert sh.first_phrase_start(3, nt, lt, 10) == 941
    pos0, nl = [0, 1000, 2000, 3000], [1000, 1000, 1000, 1000]
    req = sh.head_requests(pos0, nl, nt, lt, 10, halo=16)
    assert req[0] == (0, 0) and req[1][0] == req[1][1] and req[2][0] == req[2][1]   # no phrases there
    assert req[3] == (941, 2984)
    ops = sh.transfers(req, pos0, nl)
    assert ops == [(0, 3, 941, 1000), (1, 3, 1000, 2000), (2, 3, 2000, 2984)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, text, cuts, w, p, mode, q, front=4096, shared=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        from mock_backend import MockBackend
        pkg = load_package()
        job = pkg.shards.ShardedParser(None, world, rank, backend=MockBackend(), mode=mode, halo=64, front=front)
        if shared is not None:        # the peer-memory exchange path, over shared host memory
            from mock_backend import SharedMemExchange
            job._peer = SharedMemExchange(shared, rank, world)
        shard = torch.from_numpy(np.frombuffer(text[cuts[rank]:cuts[rank + 1]], np.uint8).copy())
        job.set_text(shard)
        st = job.parse_device(w, p, sai=True)
        files = job.gather_files()
        if rank == 0:
            q.put((files, st))
    finally:
        dist.destroy_process_group()


def run_sharded(text, cuts, w, p, mode="replicate", front=4096, peer=False):
    world = len(cuts) - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shared = None
    if peer:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from mock_backend import SharedMemExchange
        shared = SharedMemExchange.allocate(world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, text, cuts, w, p, mode, q, front, shared))
             for r in range(world)]
    for pr in procs:
        pr.start()
    files, st = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    return files, st


@pytest.mark.parametrize("world,w,p,mode", [(2, 10, 100, "replicate"), (3, 10, 100, "replicate"),
                                             (2, 4, 10, "replicate"), (3, 6, 50, "replicate"),
                                             (2, 10, 100, "partition"), (3, 10, 100, "partition"),
                                             (3, 4, 10, "partition")])
def test_sharded_protocol_matches_oracle(pkg, world, w, p, mode):
    text = pkg.synth.pangenome_text(1500, 6, 7).numpy().tobytes()
    n = len(text)
    cuts = [0] + [n * k // world + 3 * k for k in range(1, world)] + [n]
    files, st = run_sharded(text, cuts, w, p, mode)
    want = orc.parse(text, w, p)
    for ext in FILES:
        assert files[ext] == getattr(want, ext), f".{ext} differs (world {world})"
    assert st["n_phrases"] == want.n_phrases and st["n_distinct"] == want.n_distinct


def test_sharded_seam_without_triggers(pkg):
    """A shard with no trigger at all (one long run): the phrase straddles two seams and its head
    is fetched from the ranks before the halo."""
    a = pkg.synth.random_dna(3000, 3).numpy().tobytes()
    text = a + b"N" * 2500 + a[:2000]
    cuts = [0, 3200, 5000, len(text)]          # shard 1 lies inside the N run
    files, _ = run_sharded(text, cuts, 10, 100, "partition")
    want = orc.parse(text, 10, 100)
    for ext in FILES:
        assert files[ext] == getattr(want, ext), f".{ext} differs"


def test_sharded_head_longer_than_the_reserved_front(pkg):
    """The phrase straddling into the last shards starts far more bytes before them than the
    buffer reserves in front (front = 256 here; 1 MB in production): the buffer grows instead of
    the parse failing, and two whole shards lie inside the run."""
    a = pkg.synth.random_dna(2000, 4).numpy().tobytes()
    text = a + b"N" * 9000 + a[:1500]
    cuts = [0, 2500, 5000, 8000, len(text)]     # shards 1 and 2 lie inside the N run
    for mode in ("partition", "replicate"):
        files, _ = run_sharded(text, cuts, 10, 100, mode, front=256)
        want = orc.parse(text, 10, 100)
        for ext in FILES:
            assert files[ext] == getattr(want, ext), f".{ext} differs ({mode})"


@pytest.mark.parametrize("world,w,p", [(2, 10, 100), (3, 10, 100), (4, 6, 50)])
def test_sharded_peer_memory_exchange_offsets(pkg, world, w, p):
    """mode "partition" through the peer-memory branch of _merge_partitioned (the default on the
    GPUs): every rank stores its routed words straight into the owners' buffers behind the
    segments of the lower ranks, and the ranks travel back the same way."""
    text = pkg.synth.pangenome_text(1200, 7, 11).numpy().tobytes()
    n = len(text)
    cuts = [0] + [n * k // world - 5 * k for k in range(1, world)] + [n]
    files, st = run_sharded(text, cuts, w, p, "partition", peer=True)
    want = orc.parse(text, w, p)
    for ext in FILES:
        assert files[ext] == getattr(want, ext), f".{ext} differs (world {world})"
    assert st["n_phrases"] == want.n_phrases and st["n_distinct"] == want.n_distinct
