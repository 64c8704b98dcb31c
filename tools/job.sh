cd $GRAFT_REPO_ROOT
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/run_sharded.py 8192 28672 2048 --check --reps 2 2>&1 | grep "^{" | tail -1
done
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tail -1 | cut -c1-1300
done
