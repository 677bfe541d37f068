#!/bin/bash
# Round-2 starter for the multi-GPU experiments, ONE gpurun call:
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/mgpu_shot.sh 2'      (then --gpus 4, then --gpus 8 for the bench lines)
# 1. the gated parity tests: dynamic schedule on slabs, peer-memory halos (N ranks must equal one GPU bit for bit);
# 2. the weak-scaling bench line of N GPUs for: default (NCCL per-pass exchange), FXB_P2P=1, FXB_TAIL=1, both.
# Everything lands in gpurun_out/mgpu_<N>_*.
N=${1:-2}
mkdir -p gpurun_out
export FXB_TEST_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/mgpu_${N}_tests.log 2>&1
tail -5 gpurun_out/mgpu_${N}_tests.log
unset FXB_TEST_EXPERIMENTAL
run() {  # label, env...
    label=$1; shift
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-c3 --no-experiments --no-export-e2e \
        > gpurun_out/mgpu_${N}_${label}.log 2>&1
    tail -1 gpurun_out/mgpu_${N}_${label}.log | cut -c1-400
}
run default FXB_NONE=0
run p2p FXB_P2P=1
run tail FXB_TAIL=1
run p2p_tail FXB_P2P=1 FXB_TAIL=1
