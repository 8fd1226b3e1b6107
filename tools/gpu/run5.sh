cd $GRAFT_REPO_ROOT
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_forward_full python tools/profile_forward.py 3 > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
ls -la gpurun_out/
