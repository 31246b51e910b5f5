mkdir -p gpurun_out/r2
python -m pytest tests/test_multi_gpu.py tests/test_ingest_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > gpurun_out/r2/t2b.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2/t2b.log') if x.startswith('{')]
d=json.loads(l[-1]); print(round(d['value'],1), d['t2_file_to_files'])
PY
