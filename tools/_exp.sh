mkdir -p gpurun_out/r2
PFPB200_SPLIT_K3=1 python -m pytest tests/test_fullsize_gpu.py -x -q -k config2 2>&1 | tail -40 | cut -c1-400 > gpurun_out/r2/t_full_split.log
