mkdir -p gpurun_out/r2b
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b/smoke.log 2>&1; tail -3 gpurun_out/r2b/smoke.log
python bench.py > gpurun_out/r2b/bench_default.log 2> gpurun_out/r2b/bench_default.err; tail -c 300 gpurun_out/r2b/bench_default.err
python - <<PY
import json
for l in open("gpurun_out/r2b/bench_default.log"):
    if l.startswith("{"):
        j=json.loads(l); print(j["ms_per_step"], j["value"], j["roofline"]["frac"], j["e2e"]["value"], j["parity_check"]["ok"], j.get("pipeline_to_bwt"))
PY
python -m pytest tests -x -q -m gpu > gpurun_out/r2b/pytest_gpu3.log 2>&1; tail -5 gpurun_out/r2b/pytest_gpu3.log
