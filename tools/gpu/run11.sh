cd $GRAFT_REPO_ROOT
python tools/probes/geo_grid_probe.py 2>&1 | tail -18
