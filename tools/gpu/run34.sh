cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "fps" 2>&1 | grep -B5 -A25 "Error\|assert" | head -80
