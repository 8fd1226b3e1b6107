cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_res_kernel" --launch-skip 9 -c 3 -o gpurun_out/fp1 -f python tools/probes/fp1_real_timeline.py > gpurun_out/fp1_ncu.log 2>&1
ncu -i gpurun_out/fp1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:60], r[h.index('gpu__time_duration.sum')], r[h.index('launch__grid_size')])
"
ls -la gpurun_out/fp1.ncu-rep
