cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "statistics or hessian or presets or config" > gpurun_out/pytest_s3q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s3q.log
timeout 300 python tools/microbench.py xtx 2>&1 | tail -4
timeout 120 python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from sleekit_b200 import ops
for S,n in ((2048,768),(2048,1024),(512,768),(4096,200),(2048,3072)):
    X=torch.randn(S,n,device='cuda')
    H=torch.zeros(n,n,device='cuda'); m=torch.zeros(n,device='cuda')
    ops.hessian_accum(X,H,m,0.0,S)
    ref=(X.double().T@X.double()/S)
    err=float((H.double()-ref).abs().max()/ref.abs().max())
    sym=bool(torch.equal(H,H.T))
    print(S,n,'rel err',err,'symmetric',sym)
PY
