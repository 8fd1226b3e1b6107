cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "e2e_labels", "sequential", "roofline", "clocks", "gpu_launches", "train", "dp_check"):
    print(k, json.dumps(r.get(k))[:400])
for k in r["roofline_all"]["kernels"]:
    print(f"{k['launch_ms']*1e3:8.1f} us frac {k['frac']:.3f} {k['name'][:70]}")
PY
