set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -15
timeout 600 python tools/pipeline_sweep.py --steps 40 > gpurun_out/r02_sweep1.txt 2>&1; tail -30 gpurun_out/r02_sweep1.txt
