cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2o.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_s2o.log
timeout 300 python tools/microbench.py sweep > gpurun_out/micro_s2o.log 2>&1; tail -5 gpurun_out/micro_s2o.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_s2o.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/bench_s2o.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["ms_per_step"], d["serial_phases_ms_per_step"], d["layer_error_mean"], d["gpu_launches"])
PY
tail -3 gpurun_out/bench_s2o.log | cut -c1-300
