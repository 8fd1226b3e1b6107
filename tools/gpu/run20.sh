cd $GRAFT_REPO_ROOT
timeout 600 python tools/probes/e2e_probe.py 2>&1 | tail -20
