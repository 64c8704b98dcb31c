cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3f.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_s3f.log
timeout 900 python tools/config_times.py 2>&1 | grep "C3\|C4" | cut -c1-600
