#!/bin/bash
# Usage: bash tools/run_multi.sh N "c2 c3 ..."   -- one bench line per config at N GPUs (torchrun), plus the reference arm of c2
N=$1; shift
CONFIGS=${1:-"c2 c3 c4 c5 reinhard"}
PORT=29700
for c in $CONFIGS; do
  PORT=$((PORT+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --config $c --no-extras --steps 20 --warmup 5 > gpurun_out/r02_bench_${c}_n$N.json 2> gpurun_out/r02_bench_${c}_n$N.err
  echo "$c N=$N rc=$? $(head -c 200 gpurun_out/r02_bench_${c}_n$N.json)"
done
PORT=$((PORT+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm_n$N.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) tools/check_peers.py > gpurun_out/r02_check_peers_n$N.log 2>&1; tail -1 gpurun_out/r02_check_peers_n$N.log
