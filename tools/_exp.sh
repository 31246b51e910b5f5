mkdir -p gpurun_out/r2
python bench.py --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu-baseline --no-t2 > gpurun_out/r2/plain_before_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(kr_|phrase_|table_|rank_|rs_|pool_|dict_|remap_|scan_|alpha_|groups_|tile_|special_|set_u64|first_|merge_|route_|gather_|unpack_|verify_|ranks_|dna_)' --csv --log-file gpurun_out/r2/launches_4GB.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-parity --no-cpu-baseline --no-t2 > gpurun_out/r2/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'kr_scan_dna_k|phrase_stream_k|table_insert_k|rank_window_k|rank_warp_k|rank_lcp_k|pool_copy_k|dict_copy_k|rs_scatter_k' --launch-skip 14 -c 14 -o gpurun_out/r2/prof_r02 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-parity --no-cpu-baseline --no-t2 > gpurun_out/r2/ncu_full.log 2>&1
ls -la gpurun_out/r2/prof_r02.ncu-rep
