cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/run_sharded.py 8192 28672 2048 --check --reps 2 > gpurun_out/r2_shard${N}_final.log 2>&1; tail -1 gpurun_out/r2_shard${N}_final.log | cut -c1-900
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_final_n${N}.log 2>&1
tail -c 300 gpurun_out/r2_bench_final_n${N}.log
fi
