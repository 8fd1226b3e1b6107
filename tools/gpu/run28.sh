cd $GRAFT_REPO_ROOT
python tools/probes/res_timeline.py sa1 2>&1 | grep -v Warn | head -4
python tools/probes/res_timeline.py sa1 512 2>&1 | grep -v Warn | head -4
python tools/probes/res_timeline.py sa2 2>&1 | grep -v Warn | head -3
python tools/probes/res_timeline.py sa2 512 2>&1 | grep -v Warn | head -3
