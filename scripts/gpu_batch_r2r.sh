ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2r_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-membrane > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gate_rows|k_cheap_flat|k_patch_flat|k_combine_flat" -s 4 -c 4 -o gpurun_out/r2r_pipeline -f python scripts/profile_step.py everyone 1 > gpurun_out/ncu_pipe_r2r.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py rounds > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
tail -n 3 gpurun_out/r2_sanitizer_memcheck.txt gpurun_out/r2_sanitizer_racecheck.txt; wc -l gpurun_out/launches_r2r_bench.csv
