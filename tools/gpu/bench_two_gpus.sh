cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --impl reference > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -c 300 gpurun_out/bench_ref_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 400 gpurun_out/bench_n2.err; python - <<'PY'
import json
r = json.loads([l for l in open('gpurun_out/bench_n2.json').read().strip().splitlines() if l.startswith("{")][-1])
for k in ("value", "n_gpus", "ms_per_step", "e2e", "e2e_labels", "clocks", "gpu_launches", "train", "dp_check"):
    print(k, json.dumps(r.get(k))[:300])
PY
