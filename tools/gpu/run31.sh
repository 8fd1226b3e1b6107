cd $GRAFT_REPO_ROOT
timeout 300 python tools/probes/c1_timeline.py 2>&1 | grep -v Warn | tail -45
