cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q -k "scale" > gpurun_out/pytest_s3l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s3l.log
timeout 300 python tools/microbench.py search 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_s3l.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/bench_s3l.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print(d["ms_per_step"], d["serial_phases_ms_per_step"], d["layer_error_mean"], d["gpu_launches"])
PY
