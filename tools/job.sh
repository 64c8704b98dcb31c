set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q -k "cholesky_form" > gpurun_out/pytest_s2b.log 2>&1; echo "pytest rc=$?" 
tail -30 gpurun_out/pytest_s2b.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2b_all.log 2>&1; echo "pytest all rc=$?"
tail -5 gpurun_out/pytest_s2b_all.log
timeout 300 python tools/microbench.py sweep > gpurun_out/micro_s2b.log 2>&1; tail -20 gpurun_out/micro_s2b.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_s2b.log 2>&1; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_s2b.log
