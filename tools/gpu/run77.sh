cd $GRAFT_REPO_ROOT
for bits in 1024 0 1024; do
echo "ENGINE_BITS=$bits"
PN12_MLP_ENGINE_BITS=$bits timeout 300 python tools/probes/fp1_real_timeline.py 2>&1 | grep -v Warn | grep "round [4-5]" | cut -c1-330
PN12_MLP_ENGINE_BITS=$bits timeout 300 python tools/kernel_rooflines.py 2>/dev/null | python -c "
import json,sys
r=json.load(sys.stdin)
for k in r['kernels']:
    if 'fp1 +' in k['name'] or 'sa2' in k['name']: print(f\"{k['launch_ms']*1e3:8.1f} us  {k['name'][:50]}\")
"
PN12_MLP_ENGINE_BITS=$bits timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {}}' 2>&1 | grep "depth"
done
