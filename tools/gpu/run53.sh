cd $GRAFT_REPO_ROOT
PN12_NN_VARIANT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"three_nn_blocks|nn_blocks_build" -c 4 -o gpurun_out/nn1 -f python tools/probes/nn1_ab.py > gpurun_out/nn1_ncu.log 2>&1
ncu -i gpurun_out/nn1.ncu-rep --page raw --csv > gpurun_out/nn1_raw.csv 2>/dev/null
ls -la gpurun_out/nn1*
