mkdir -p gpurun_out/r2
python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu.py tests/test_shards_gpu.py -x -q 2>&1 | tail -2
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-t2"
$B > gpurun_out/r2/v8_100.log 2>&1
PFPB200_POOL_BY_WORD=1 $B > gpurun_out/r2/v8_100_byword.log 2>&1
$B --workload random --base-len 80000000 --haplotypes 100 --steps 3 --warmup 2 > gpurun_out/r2/v8_rand.log 2>&1
for f in v8_100 v8_100_byword v8_rand; do grep -h '"value"' gpurun_out/r2/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items() if k in ('ms_rank','ms_dedup','ms_hash','ms_dict','ms_scan')})
"; done
