cd $GRAFT_REPO_ROOT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "three_nn or nn_blocks or ball or fps_few_wide or fps_every_cluster or dynamic_tiles or labels or small_n" > gpurun_out/r02_sanitizer.log 2>&1
echo "exit $?"; tail -5 gpurun_out/r02_sanitizer.log; grep -c "ERROR SUMMARY" gpurun_out/r02_sanitizer.log; grep "ERROR SUMMARY" gpurun_out/r02_sanitizer.log | sort | uniq -c | head
