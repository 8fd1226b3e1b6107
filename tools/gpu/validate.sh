cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; tail -c 600 gpurun_out/bench_ref_n1.json
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "e2e_labels", "sequential", "roofline", "clocks", "gpu_launches", "train", "dp_check"):
    print(k, json.dumps(r.get(k))[:300])
PY
