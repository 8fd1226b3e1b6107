cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "pointnet_seg or heads or pointnet_cls or dense or train_mode or pointnet" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py 2>&1 | grep "C1"
