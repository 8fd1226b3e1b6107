cd $GRAFT_REPO_ROOT
PN12_FP1_ORDER=1 timeout 900 python -m pytest tests -x -q -m gpu -k "semseg or golden or pipelined" 2>&1 | tail -2
for o in 0 1; do
echo "FP1_ORDER=$o"
PN12_FP1_ORDER=$o python tools/probes/fp1_real_timeline.py 2>&1 | grep -v Warn | grep "round [3-5]" | cut -c1-75,330-420
PN12_FP1_ORDER=$o python tools/kernel_rooflines.py 2>/dev/null | python -c "
import json,sys
r=json.load(sys.stdin)
for k in r['kernels']:
    if 'fp1' in k['name']: print(f\"{k['launch_ms']*1e3:8.1f} us  {k['name']}\")
"
PN12_FP1_ORDER=$o timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {}}' 2>&1 | grep "depth"
done
