mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/b8_weak2.log 2> gpurun_out/r2/b8_weak2.err
$TR bench.py --gpus 8 --steps 5 --warmup 2 --base-len 64000000 --haplotypes 125 --no-e2e > gpurun_out/r2/c3_n8b.log 2> gpurun_out/r2/c3_n8b.err
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/r2/t8_multi2.log
# the C path on 8 GPUs: 8 GB of plain text and the same as FASTA through gpuscan.x -g 0..7, twice each
python - > gpurun_out/r2/cli8b.log 2>&1 <<'PY'
import os, subprocess, sys, time, hashlib
sys.path.insert(0, '.')
from __graft_entry__ import load_package
pkg = load_package()
os.makedirs("/tmp/cli8", exist_ok=True)
recs = list(pkg.synth.pangenome_records(40_000_000, 200, 2, device="cuda"))
with open("/tmp/cli8/a.txt", "wb") as f, open("/tmp/cli8/a.fa", "wb") as g:
    for k, r in enumerate(recs):
        r = r.cpu().numpy(); r.tofile(f); pkg.synth.to_fasta_np(r, f"hap{k}").tofile(g)
del recs
def run(cmd):
    t0 = time.time(); r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True); return time.time() - t0, r
for path, extra in (("/tmp/cli8/a.txt", []), ("/tmp/cli8/a.fa", ["-f"])):
    for rep in range(2):
        dt, r = run([pkg.pfp.CLI_PATH, path, "-w", "10", "-p", "100", "-s", "-v", "-g", "0,1,2,3,4,5,6,7"] + extra)
        print(path, "rep", rep, round(dt, 2), "s rc", r.returncode, "size", os.path.getsize(path))
    print(r.stdout[-2600:])
dt1, r1 = run([pkg.pfp.CLI_PATH, "/tmp/cli8/a.txt", "-w", "10", "-p", "100", "-s"])
sha8 = {e: hashlib.sha256(open(f"/tmp/cli8/a.fa.{e}", "rb").read()).hexdigest()[:16] for e in ("dict", "occ", "parse", "last", "sai")}
sha1 = {e: hashlib.sha256(open(f"/tmp/cli8/a.txt.{e}", "rb").read()).hexdigest()[:16] for e in ("dict", "occ", "parse", "last", "sai")}
print("1 GPU plain:", round(dt1, 2), "s"); print(r1.stdout[-500:])
print("8-GPU FASTA files == 1-GPU plain files:", sha8 == sha1, sha8)
PY
tail -2 gpurun_out/r2/t8_multi2.log; tail -4 gpurun_out/r2/cli8b.log; tail -c 200 gpurun_out/r2/b8_weak2.err
