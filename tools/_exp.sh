python -m pytest tests/test_multi_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-t2"
$B 2>&1 | grep '"value"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items()})
"
