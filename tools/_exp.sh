mkdir -p gpurun_out/r2
python -m pytest tests/test_parity_gpu.py -x -q -k "rank or low_complexity or long_phrases or pangenome" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(rank_|rs_|alpha_|groups_|dict_|remap_)' --csv --log-file gpurun_out/r2/launches_v6_800.csv python bench.py --steps 1 --warmup 1 --haplotypes 800 --no-e2e --no-parity --no-cpu-baseline --no-t2 > /dev/null 2>&1
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-t2"
$B > gpurun_out/r2/v6_100.log 2>&1
$B --haplotypes 800 --steps 3 --warmup 2 > gpurun_out/r2/v6_800.log 2>&1
for f in v6_100 v6_800; do grep -h '"value"' gpurun_out/r2/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items() if k in ('ms_rank','ms_dedup','ms_hash','ms_dict','ms_scan')})
"; done
