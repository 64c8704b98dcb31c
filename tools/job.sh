set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2a.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_s2a.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2a.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_s2a.log
timeout 300 python bench.py --steps 5 --warmup 3 --only big --no-e2e --no-cpu-baseline > gpurun_out/bench_s2a_big.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --only small --no-e2e --no-cpu-baseline > gpurun_out/bench_s2a_small.log 2>&1
python - <<'PY'
import json
for f in ("big","small"):
    for ln in open(f"gpurun_out/bench_s2a_{f}.log"):
        if ln.startswith("{"):
            d=json.loads(ln); print(f, d["ms_per_step"], d["serial_phases_ms_per_step"])
PY
