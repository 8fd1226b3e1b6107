cd $GRAFT_REPO_ROOT
run() {
  for i in 1 2; do
  env $1 timeout 900 python bench.py --no-train --depth $2 > gpurun_out/b.json 2> gpurun_out/b.err; tail -c 300 gpurun_out/b.err; python - "$1" $2 <<'PY'
import json, sys
r = json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1])
print(sys.argv[1], "depth", sys.argv[2], " ".join(f"{k}={r[k]['ms_per_step'] if isinstance(r[k], dict) else r[k]:.4f}" for k in ("ms_per_step", "e2e", "e2e_labels")))
PY
  done
}
run X=1 6
run PN12_FPS1=3,256,2 6
run PN12_FPS1=3,256,2 8
run PN12_FPS1=2,256,2 8
run PN12_FPS1=2,256,2 10
