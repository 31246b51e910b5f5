#!/usr/bin/env python3
"""Probe: peer-memory all-to-all over NVLink with torch symmetric memory vs NCCL grouped send/recv.
torchrun --nproc-per-node G tools/symm_probe.py [MB per peer]"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = mb << 20
send = torch.full((world * n,), rank + 1, dtype=torch.uint8, device=dev)
# --- NCCL grouped P2P
recv = torch.empty(world * n, dtype=torch.uint8, device=dev)
def nccl_a2a():
    ops = []
    for q in range(world):
        if q == rank:
            recv[q * n:(q + 1) * n] = send[q * n:(q + 1) * n]; continue
        ops.append(dist.P2POp(dist.isend, send[q * n:(q + 1) * n], q))
        ops.append(dist.P2POp(dist.irecv, recv[q * n:(q + 1) * n], q))
    for r in dist.batch_isend_irecv(ops): r.wait()
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_nccl = timeit(nccl_a2a)
t_a2a = timeit(lambda: dist.all_to_all_single(recv, send))
# --- symmetric memory: every rank stores straight into the owners' buffers
try:
    buf = symm_mem.empty(world * n, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    def symm_a2a(multi_stream=True):
        cur = torch.cuda.current_stream()
        hdl.barrier()                       # receivers are done with the previous contents
        for i in range(world):
            q = (rank + i) % world
            dst = hdl.get_buffer(q, (world * n,), torch.uint8)[rank * n:(rank + 1) * n]
            if multi_stream:
                s = streams[i]; s.wait_stream(cur)
                with torch.cuda.stream(s): dst.copy_(send[q * n:(q + 1) * n], non_blocking=True)
            else:
                dst.copy_(send[q * n:(q + 1) * n], non_blocking=True)
        if multi_stream:
            for s in streams: cur.wait_stream(s)
        hdl.barrier()                       # everything addressed to me has landed
    t_s1 = timeit(lambda: symm_a2a(False))
    t_sm = timeit(lambda: symm_a2a(True))
    ok = all(int(buf[q * n]) == q + 1 and int(buf[(q + 1) * n - 1]) == q + 1 for q in range(world))
    msg = f"symm single-stream {t_s1:.3f} ms, multi-stream {t_sm:.3f} ms, data ok={ok}"
except Exception as e:  # noqa: BLE001
    msg = f"symmetric memory unavailable: {type(e).__name__}: {e}"
gb = (world - 1) * n / 1e9
if rank == 0:
    print(f"world {world}, {mb} MB per peer ({gb:.2f} GB leave each rank): NCCL grouped p2p {t_nccl:.3f} ms "
          f"({gb / t_nccl * 1e3:.0f} GB/s), all_to_all_single {t_a2a:.3f} ms ({gb / t_a2a * 1e3:.0f} GB/s); {msg}", flush=True)
dist.destroy_process_group()
