#!/bin/bash
# What the N-GPU step loses against one GPU: rank-to-rank variance (no exchange at all) vs the NCCL kernels sharing
# SMs / HBM with the backward (channel count varied).  N = $1.
N=${1:-4}
P=29600
run() {  # label, env...
  P=$((P+1)); local label=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --config ${CFG:-2} --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', d['n_gpus'], round(d['ms_per_step'], 3), d.get('grad_sync') and d['grad_sync']['exposed_wait_ms_last_step'], d['clocks']['sm_mhz'])
"
}
run default A=1
run no_sync B200MM_NO_SYNC=1
run max_ctas_2 NCCL_MAX_CTAS=2
run max_ctas_4 NCCL_MAX_CTAS=4
run max_ctas_8 NCCL_MAX_CTAS=8
run max_ctas_16 NCCL_MAX_CTAS=16
