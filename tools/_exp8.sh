mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/b8_weak.log 2> gpurun_out/r2/b8_weak.err
$TR bench.py --gpus 8 --steps 3 --warmup 2 --base-len 64000000 --haplotypes 125 --no-e2e > gpurun_out/r2/c3_n8.log 2> gpurun_out/r2/c3_n8.err
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r2/t8_multi.log
# the C path on 8 GPUs: 2 GB of plain text through gpuscan.x -g 0..7 against the single-GPU files
python - > gpurun_out/r2/cli8.log 2>&1 <<'PY'
import os, subprocess, sys, time, hashlib
sys.path.insert(0, '.')
from __graft_entry__ import load_package
pkg = load_package()
t = pkg.synth.pangenome_text(40_000_000, 50, 2, device="cuda").cpu().numpy()
os.makedirs("/tmp/cli8", exist_ok=True)
for name in ("a", "b"):
    t.tofile(f"/tmp/cli8/{name}.txt")
def run(cmd):
    t0 = time.time(); r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True); return time.time() - t0, r
dt8, r8 = run([pkg.pfp.CLI_PATH, "/tmp/cli8/a.txt", "-w", "10", "-p", "100", "-s", "-v", "-g", "0,1,2,3,4,5,6,7"])
dt1, r1 = run([pkg.pfp.CLI_PATH, "/tmp/cli8/b.txt", "-w", "10", "-p", "100", "-s"])
print("8 GPUs:", round(dt8, 2), "s rc", r8.returncode); print(r8.stdout[-2500:])
print("1 GPU:", round(dt1, 2), "s rc", r1.returncode); print(r1.stdout[-600:])
same = all(hashlib.sha256(open(f"/tmp/cli8/a.txt.{e}", "rb").read()).digest() == hashlib.sha256(open(f"/tmp/cli8/b.txt.{e}", "rb").read()).digest() for e in ("dict", "occ", "parse", "last", "sai"))
print("files identical:", same)
PY
tail -2 gpurun_out/r2/t8_multi.log; tail -5 gpurun_out/r2/cli8.log; tail -c 300 gpurun_out/r2/b8_weak.err; tail -c 300 gpurun_out/r2/c3_n8.err
