cd $GRAFT_REPO_ROOT
python tools/probes/fp1_real_timeline.py > gpurun_out/r02_fp1_timeline.txt 2>&1
cat gpurun_out/r02_fp1_timeline.txt
