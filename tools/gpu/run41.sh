cd $GRAFT_REPO_ROOT
timeout 600 python tools/probes/e2e_hosttime.py 2>&1 | grep -v Warn
