cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_geometry_full python tools/profile_geometry.py > gpurun_out/r02_ncu_geo.log 2>&1
tail -2 gpurun_out/r02_ncu_geo.log
