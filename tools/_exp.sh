mkdir -p gpurun_out/r2
python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2/t_lcp.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-t2"
$B > gpurun_out/r2/lcp_100.log 2>&1
PFPB200_RANK_CHUNK_PASSES=1 $B > gpurun_out/r2/chk_100.log 2>&1
$B --haplotypes 800 --steps 3 --warmup 2 > gpurun_out/r2/lcp_800.log 2>&1
PFPB200_RANK_CHUNK_PASSES=1 $B --haplotypes 800 --steps 3 --warmup 2 > gpurun_out/r2/chk_800.log 2>&1
python tools/fullsize_check.py --config sweep > gpurun_out/r2/sweep_lcp.jsonl 2> gpurun_out/r2/sweep_lcp.err
cat gpurun_out/r2/t_lcp.log
for f in lcp_100 chk_100 lcp_800 chk_800; do grep -h '"value"' gpurun_out/r2/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items() if k in ('ms_rank','ms_dedup','ms_hash','ms_dict')})
"; done
python - <<'PY'
import json
for l in open('gpurun_out/r2/sweep_lcp.jsonl'):
    d = json.loads(l); print(d['w'], d['p'], d['GBps'], d['stages_ms']['rank'], d['ok'], d.get('sha256_vs_reference'))
PY
