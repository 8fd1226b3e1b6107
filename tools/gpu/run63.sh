cd $GRAFT_REPO_ROOT
python tools/probes/bq1_probe.py 2>&1 | grep "ball query"
