cd $GRAFT_REPO_ROOT
for k in "fps_async_kernel<8, 48" "three_nn_blocks_kernel" "mlp_tc_res_kernel<2, 8>" "mlp_tc_res_kernel<1, 8>"; do
  tag=$(echo "$k" | tr -c 'a-z0-9_' '_' | cut -c1-24)
  timeout 300 ncu --set full --clock-control none -k regex:"$k" --launch-skip 2 -c 1 -o gpurun_out/r02_fin_$tag -f python tools/kernel_rooflines.py --fps1 2,256,2 > gpurun_out/r02_fin_$tag.log 2>&1
  ncu -i gpurun_out/r02_fin_$tag.ncu-rep --page raw --csv > gpurun_out/r02_fin_$tag.csv 2>/dev/null
done
ls -la gpurun_out/r02_fin_*.csv
