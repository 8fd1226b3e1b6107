cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do
timeout 900 python bench.py --no-train > gpurun_out/b.json 2> gpurun_out/b.err; tail -c 300 gpurun_out/b.err; python - <<'PY'
import json, sys
r = json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1])
print(" ".join(f"{k}={r[k]['ms_per_step'] if isinstance(r[k], dict) else r[k]:.4f}" for k in ("ms_per_step", "e2e", "e2e_labels")), r["roofline"]["frac"], r["gpu_launches"])
PY
done
