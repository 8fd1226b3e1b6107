cd $GRAFT_REPO_ROOT
E2E_DEPTHS=6,8,10,12 E2E_SLICES=1 timeout 600 python tools/probes/e2e_probe.py 2>&1 | grep -v Warn
