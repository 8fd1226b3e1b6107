cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "fp or semseg or partseg or golden or heads or resident" 2>&1 | tail -2
python tools/probes/fp1_real_timeline.py 2>&1 | grep -v Warn | grep "round [2-5]" | cut -c1-120,330-420
python tools/kernel_rooflines.py 2>/dev/null | python -c "
import json,sys
r=json.load(sys.stdin)
for k in r['kernels']:
    if 'fp1' in k['name']: print(f\"{k['launch_ms']*1e3:8.1f} us  {k['name']}\")
"
timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {}}' 2>&1 | grep "depth"
