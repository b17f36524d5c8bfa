#!/bin/bash
# Multi-GPU correctness + bench on one box: tools/scale_run.sh N [out-prefix]. Every step runs under its own timeout.
N=${1:-2}; OUT=${2:-gpurun_out/scale_n$N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
{
if [ "$N" -le 2 ]; then MODELS="tiny-llama31 tiny-qwen3 tiny-qwen3-moe"; elif [ "$N" -le 4 ]; then MODELS="tiny-llama-tp8 tiny-llama-kvrep tiny-qwen3-moe"; else MODELS="tiny-llama-tp8 tiny-llama-kvrep"; fi
for m in $MODELS; do
  if [ -z "$QUICK" ]; then echo "== tp_check $m (multi-kernel decode)"; timeout 240 $TR tools/tp_check.py $m 2>&1 | grep -v "^W\|^\[W\|warn" | tail -4; fi
  echo "== tp_check $m (persistent kernel)"; DL_MEGA=1 timeout 240 $TR tools/tp_check.py $m 2>&1 | grep -v "^W\|^\[W\|warn" | tail -4
done
echo "== bench.py --gpus $N"; timeout 400 $TR bench.py --gpus $N --steps 64 --warmup 4 2>&1 | grep "^{" | tail -1
} > $OUT.log 2>&1
cat $OUT.log
