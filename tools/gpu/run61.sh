cd $GRAFT_REPO_ROOT
python tools/probes/bq1_probe.py 2>&1 | grep "ball query"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ball_query_grid_kernel" --launch-skip 3 -c 1 -o gpurun_out/bq1 -f python tools/probes/bq1_probe.py > gpurun_out/bq1_ncu.log 2>&1
ls -la gpurun_out/bq1.ncu-rep
