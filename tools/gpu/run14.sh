cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_fps_sorted python tools/probes/fps_sorted_ncu.py > gpurun_out/ncu_fps.log 2>&1
tail -2 gpurun_out/ncu_fps.log
