mkdir -p gpurun_out/r2b
python -m pytest tests -x -q -m gpu > gpurun_out/r2b/pytest_gpu.log 2>&1; tail -5 gpurun_out/r2b/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench_full.log 2> gpurun_out/r2b/bench_full.err; tail -c 400 gpurun_out/r2b/bench_full.err
Q="--no-cpu-baseline --no-ref-full --no-e2e --no-t2 --no-parity --steps 2 --warmup 1"
python bench.py $Q > gpurun_out/r2b/plain_before_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b/launches.csv python bench.py $Q > gpurun_out/r2b/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:kr_scan_ivf_k --launch-skip 1 --launch-count 1 -o gpurun_out/r2b/prof_ivf python bench.py $Q > gpurun_out/r2b/ncu_ivf_full.log 2>&1
ls -la gpurun_out/r2b
