cd $GRAFT_REPO_ROOT
timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {"PN12_WHATIF": "fps1"}};{"depth": 10, "env": {"PN12_WHATIF": "fp1"}};{"depth": 10, "env": {"PN12_WHATIF": "sa1"}};{"depth": 10, "env": {"PN12_WHATIF": "nn1"}};{"depth": 10, "env": {"PN12_WHATIF": "bq1"}};{"depth": 10, "env": {"PN12_WHATIF": "sa2"}};{"depth": 10, "env": {"PN12_WHATIF": "fps1,fp1,sa1,nn1,bq1,sa2"}};{"depth": 6, "env": {}};{"depth": 8, "env": {}};{"depth": 12, "env": {}}' 2>&1 | grep depth > gpurun_out/whatif_d10.txt; cat gpurun_out/whatif_d10.txt
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; tail -c 200 gpurun_out/bench_ref_n1.json
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "e2e_labels", "sequential", "roofline", "clocks", "gpu_launches", "train", "dp_check"):
    print(k, json.dumps(r.get(k))[:260])
for k in r["roofline_all"]["kernels"]:
    print(f"{k['launch_ms']*1e3:8.1f} us frac {k['frac']:.3f} {k['name'][:60]}")
PY
