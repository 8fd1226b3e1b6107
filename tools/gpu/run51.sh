cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "three_nn or nn_blocks or interpol or fp or semseg" 2>&1 | tail -3
python tools/kernel_rooflines.py 2>/dev/null | python -c "
import json,sys
r=json.load(sys.stdin)
for k in r['kernels']:
    print(f\"{k['launch_ms']*1e3:8.1f} us  {k['name']}\")
print('sum', r['sum_ms'])
"
timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {}}' 2>&1 | grep depth
