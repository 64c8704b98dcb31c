cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -s > gpurun_out/pytest_cfg.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_cfg.log
