cd $GRAFT_REPO_ROOT
timeout 600 python tools/run_sharded.py 8192 28672 2048 --check --reps 2 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py 8192 28672 2048 --check --reps 2 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500
