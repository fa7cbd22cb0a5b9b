cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_ceres_shim.py -x -q -m gpu 2>&1 | tail -15
