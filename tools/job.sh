cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -s 2>&1 | grep "config\|passed\|failed" | cut -c1-200
timeout 900 python tools/config_times.py 2>&1 | grep "C4\|C5" | cut -c1-560
