cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -3 gpurun_out/r02_bench_n8.err
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/r02_topo.txt
