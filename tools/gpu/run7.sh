cd $GRAFT_REPO_ROOT
python tools/probes/fp1_phase_probe.py 2>&1 | tail -12
