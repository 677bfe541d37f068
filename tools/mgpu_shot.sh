#!/bin/bash
# N-GPU bench lines (and optionally the parity tests): gpurun --gpus N --timeout 1200 -- 'bash tools/mgpu_shot.sh N tag [tests] [backends...]'
N=$1; tag=${2:-r2}; shift 2
out=gpurun_out; mkdir -p $out
if [ "$1" = tests ]; then
  shift
  timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | grep -v "^W1017" | tail -5
fi
for b in "${@:-fused}"; do
  extra=""; name=$b
  case $b in
    fused) extra="--halo fused";;
    peer) extra="--halo peer";;
    nccl) extra="--halo nccl";;
    peer_g4) extra="--halo peer --jacobi-group 4 --fuse-t 2";;
    strong) extra="--halo fused --scaling strong";;
  esac
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
      bench.py --gpus $N --steps 30 --warmup 5 $extra > $out/${tag}_bench_${N}gpu_${name}.json 2> $out/${tag}_bench_${N}gpu_${name}.err
  echo "== $name rc=$?"
  grep '^{' $out/${tag}_bench_${N}gpu_${name}.json | python tools/benchsum.py | cut -c1-700
  grep -i "error\|Traceback" $out/${tag}_bench_${N}gpu_${name}.err | head -5
done
