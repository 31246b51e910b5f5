mkdir -p gpurun_out/r2
python bench.py --gpus 1 --steps 3 --warmup 2 --base-len 64000000 --haplotypes 1000 --no-e2e --no-t2 --no-cpu-baseline > gpurun_out/r2/c3_n1.log 2> gpurun_out/r2/c3_n1.err
tail -c 600 gpurun_out/r2/c3_n1.err; nvidia-smi --query-gpu=memory.used,memory.total --format=csv
