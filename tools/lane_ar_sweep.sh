#!/bin/bash
# sweep of the lane all-reduce launch shape (torchrun, N ranks): prints the step span and the two all-reduce durations
N=${1:-2}; ROWS=${2:-1024}
for cfg in "64 128" "96 256" "148 128" "148 256" "148 512" "32 512" "74 512"; do
  set -- $cfg
  out=$(DSC_P2P_LANE_CTAS=$1 DSC_P2P_LANE_THREADS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/step_timeline.py $ROWS 2>&1)
  echo "ctas=$1 threads=$2 rows=$ROWS: $(echo "$out" | grep -o 'span [0-9.]* us') | $(echo "$out" | grep allreduce_p2p | awk '{print $1, $2}' | tr '\n' ' ') | gemm dW1 $(echo "$out" | grep gemm_tf32x3 | tail -1 | awk '{print $1, $2}')"
done
