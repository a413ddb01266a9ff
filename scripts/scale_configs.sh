#!/bin/bash
# BASELINE configs 2-5 at N GPUs of one box (N = $1), one JSON line each -> gpurun_out/bench_cfg<c>_n<N>_r02.json
N=${1:-8}
P=29500
for c in 2 3 4 5; do
  P=$((P+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --config $c --steps 10 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_cfg${c}_n${N}_r02.json 2> gpurun_out/bench_cfg${c}_n${N}_r02.err
  echo "config $c rc=$?"; tail -c 600 gpurun_out/bench_cfg${c}_n${N}_r02.json; tail -3 gpurun_out/bench_cfg${c}_n${N}_r02.err
done
