mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
python -m pytest tests/test_multi_gpu.py tests/test_shards_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/r2/t2gpu_c.log
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/b2_weak.log 2> gpurun_out/r2/b2_weak.err
$TR bench.py --gpus 2 --steps 3 --warmup 2 --base-len 64000000 --haplotypes 500 --no-e2e > gpurun_out/r2/c3_n2.log 2> gpurun_out/r2/c3_n2.err
cat gpurun_out/r2/t2gpu_c.log; tail -c 300 gpurun_out/r2/b2_weak.err; tail -c 300 gpurun_out/r2/c3_n2.err
