cd $GRAFT_REPO_ROOT
cap() {  # name skip tag
  timeout 300 ncu --set full --clock-control none -k regex:"$1" --launch-skip $2 -c 1 -o gpurun_out/r02_fin_$3 -f python tools/kernel_rooflines.py --fps1 2,256,2 > gpurun_out/r02_fin_$3.log 2>&1
  ncu -i gpurun_out/r02_fin_$3.ncu-rep --page raw --csv > gpurun_out/r02_fin_$3.csv 2>/dev/null
}
cap fps_async_kernel 2 fps1_2x8
cap mlp_tc_res_kernel 14 fp1_head
cap mlp_tc_res_kernel 8 sa2
ls -la gpurun_out/r02_fin_*.csv
