cd $GRAFT_REPO_ROOT
timeout 300 python tools/probes/e2e_matrix.py 2>&1 | grep -v Warn | head -6
for i in 1 2 3 4; do
timeout 900 python bench.py --no-train > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(" ".join(f"{k}={r[k]['ms_per_step'] if isinstance(r[k], dict) else r[k]:.4f}" for k in ("ms_per_step", "e2e", "e2e_labels")))
PY
done
timeout 600 python -m pytest tests -x -q -m gpu -k "pipelined or labels or graphed" 2>&1 | tail -2
