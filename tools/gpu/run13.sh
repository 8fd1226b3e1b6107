cd $GRAFT_REPO_ROOT
timeout 300 python tools/probes/fps_sorted_probe.py 2>&1 | tail -9
