#!/bin/bash
# Determinism / long-context check (counterpart of the reference's examples/macbeth.sh): generate a long greedy
# continuation twice — single GPU and tensor-parallel — and require identical text. Greedy decoding on this engine is
# bit-deterministic (fixed-order reductions, rank-ordered all-reduce), so the two runs must agree exactly.
#
#   MODEL=model.m TOKENIZER=tok.t N=4 STEPS=2048 bash examples/macbeth.sh
cd "$(dirname "$0")/.."
N=${N:-2}; STEPS=${STEPS:-512}
PROMPT=${PROMPT:-"Duncan. What bloody man is that? He can report, as seemeth by his plight, of the revolt the newest state."}
run() { "$@" --model "$MODEL" --tokenizer "$TOKENIZER" --buffer-float-type q80 --prompt "$PROMPT" --steps "$STEPS" --temperature 0 --seed 12345 \
        | grep "🔶 Pred" | sed 's/.*| //' | tr -d '\n'; }
A=$(run ./dllama inference)
B=$(run ./dllama inference --gpus "$N")
if [ "$A" == "$B" ] && [ -n "$A" ]; then echo "✅ identical continuation on 1 and $N GPUs (${#A} bytes)"; else echo "❌ outputs differ"; echo "1 GPU : $A"; echo "$N GPUs: $B"; exit 1; fi
