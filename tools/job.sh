cd $GRAFT_REPO_ROOT
python tools/prof_one.py 768 768 3 > gpurun_out/prof_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'chol_dag|sweep_macro|scale_search_tab|tc_gemm|chol_export|chol_gather' -s 20 -c 10 -f -o gpurun_out/prof_r1_768x768 python tools/prof_one.py 768 768 3 > gpurun_out/prof_ncu2.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/prof_ncu2.log
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from sleekit_b200 import ops
from sleekit_b200 import workloads as wl
for n in (1024, 4096, 8192, 11008):
    X = torch.randn(max(2048, n // 2), n, device='cuda')
    H = torch.zeros(n, n, device='cuda'); m = torch.zeros(n, device='cuda')
    ops.hessian_accum(X, H, m, 0.0, X.shape[0])
    damp = ops.damp_value(H, 0.01)
    order = ops.argsort(ops.order_keys(H, damp, None))
    for _ in range(2): out = ops.chol_factor(H, order, damp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = ops.chol_factor(H, order, damp); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"chol_factor n={n}: {ms:.3f} ms  {n**3/3/ms/1e9:.2f} TFLOP/s fp64  info={int(out[3].item())}")
    del X, H, out
PY
