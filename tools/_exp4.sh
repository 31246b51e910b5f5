mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543"
$TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2/b4_weak.log 2> gpurun_out/r2/b4_weak.err
$TR bench.py --gpus 4 --steps 3 --warmup 2 --base-len 64000000 --haplotypes 250 --no-e2e > gpurun_out/r2/c3_n4.log 2> gpurun_out/r2/c3_n4.err
tail -c 300 gpurun_out/r2/b4_weak.err; tail -c 300 gpurun_out/r2/c3_n4.err
