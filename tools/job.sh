cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s3p.log
timeout 600 python bench.py > gpurun_out/bench_s3p.log 2>&1; echo "bench rc=$?"; tail -c 5000 gpurun_out/bench_s3p.log
