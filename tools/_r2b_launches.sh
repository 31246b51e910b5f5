mkdir -p gpurun_out/r2b
Q="--no-cpu-baseline --no-ref-full --no-e2e --no-t2 --no-parity --steps 2 --warmup 1"
python bench.py $Q > gpurun_out/r2b/plain_before_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  -k 'regex:^(kr_|phrase_|table_|pool_|dict_|rank_|rs_|remap_|scan_|tile_|alpha_|groups_|special_|set_u64|dna_|first_|verify|fasta|bp_|up_|route|merge|uid|word|long|insert|record|meta|occ)' \
  --log-file gpurun_out/r2b/launches.csv python bench.py $Q > gpurun_out/r2b/ncu_launches.log 2>&1
grep -c kr_scan gpurun_out/r2b/launches.csv; wc -l gpurun_out/r2b/launches.csv
