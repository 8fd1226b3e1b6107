cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 4 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02_final_bench_under_ncu.json 2> gpurun_out/r02_final_bench_under_ncu.err
wc -l gpurun_out/r02_final_launches.csv
python tools/summarize_launches.py gpurun_out/r02_final_launches.csv > gpurun_out/r02_final_launches_summary.txt 2>&1; head -12 gpurun_out/r02_final_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fps_async_kernel|mlp_tc_res_kernel|three_nn_blocks_kernel|ball_query_grid_kernel" -c 16 -o gpurun_out/r02_final_full -f python tools/kernel_rooflines.py > gpurun_out/r02_final_full.log 2>&1
ncu -i gpurun_out/r02_final_full.ncu-rep --page raw --csv > gpurun_out/r02_final_ncu_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02_final*
