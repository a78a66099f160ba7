#!/bin/bash
# Multi-GPU measurement sweep on ONE box (run under gpurun --gpus N): per-step bench (weak scaling, BASELINE configs[1] per
# rank) with every gather mode, and the 1024-stream job (BASELINE configs[4], strong scaling) at N = 1, 2, 4, ... MAX,
# through bench.py (one process per GPU, torchrun) and through host/main (one process, one thread per GPU).
# usage: scripts/scale_sweep.sh <max gpus> <outdir> [streams]
MAX=${1:-8}; OUT=${2:-gpurun_out/scale_r2}; STREAMS=${3:-1024}
# PARTS: which sections to run (default all): step modes c5 main; NLIST: which N (default 1 2 4 8 up to MAX); C5G: config5 gather modes
PARTS=${PARTS:-"step modes c5 main"}; C5G=${C5G:-"nccl direct copy"}; MAING=${MAING:-"nccl direct"}
mkdir -p "$OUT"
PORT=29600
trun() { PORT=$((PORT + 1)); local n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $PORT "$@"; }
run() { local name=$1; shift; local t0=$(date +%s); "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"; local rc=$?; grep '^{' "$OUT/$name.out" | tail -1 > "$OUT/$name.json"; echo "$name rc=$rc $(( $(date +%s) - t0 ))s $(python - "$OUT/$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print("value %.1f Gb/s  without gather %s  ms/step %.4f" % (d["value"], ("%.1f" % d["value_without_gather"]) if d.get("value_without_gather") else "-", d["ms_per_step"]))
except Exception as e:
    print("no JSON line")
PY
)"; }
NS=""; for n in ${NLIST:-1 2 4 8}; do [ $n -le $MAX ] && NS="$NS $n"; done
has() { case " $PARTS " in *" $1 "*) return 0;; esac; return 1; }
echo "== per-step bench, default gather"
has step && for n in $NS; do
  if [ $n -eq 1 ]; then run step_n1 timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline
  else run step_n${n} trun $n bench.py --gpus $n --steps 200 --warmup 5; fi
done
echo "== per-step bench at N=$MAX, other gather modes"
if has modes && [ $MAX -gt 1 ]; then
  for g in ${MODES:-nccl copy direct none}; do run step_n${MAX}_$g trun $MAX bench.py --gpus $MAX --steps 200 --warmup 5 --gather $g --no-e2e; done
  case " ${MODES:-nccl} " in *" nccl "*) VIT_NCCL_MAX_CTAS=2 run step_n${MAX}_nccl_cta2 trun $MAX bench.py --gpus $MAX --steps 200 --warmup 5 --gather nccl --no-e2e;; esac
  export VIT_NCCL_MAX_CTAS=; unset VIT_NCCL_MAX_CTAS
fi
echo "== config 5: $STREAMS streams x 256 Mbit s8 / int16x2"
has c5 && for n in $(echo $NS | command tr ' ' '\n' | sort -rn); do
  for g in $C5G; do
    [ $n -eq 1 ] && [ $g != "${C5G%% *}" ] && continue
    if [ $n -eq 1 ]; then run c5_n1 timeout 300 python bench.py --workload config5 --streams $STREAMS --gather $g
    else run c5_n${n}_$g trun $n bench.py --gpus $n --workload config5 --streams $STREAMS --gather $g; fi
  done
  if [ $n -gt 1 ] && [ -n "${C5_NCCL_CTAS:-}" ]; then
    export VIT_NCCL_MAX_CTAS=$C5_NCCL_CTAS
    run c5_n${n}_nccl_cta$C5_NCCL_CTAS trun $n bench.py --gpus $n --workload config5 --streams $STREAMS --gather nccl
    unset VIT_NCCL_MAX_CTAS
  fi
done
echo "== host/main, one process, one thread per GPU"
has main && for g in $MAING; do
  timeout 300 gpu-accelerated-viterbi-decoder_b200/host/main --streams $STREAMS --gpus $MAX -n 256000000 -i s8 -m b16 --prbs --seed 1 --gather $g > "$OUT/main_n${MAX}_$g.txt" 2>&1
  echo "main_n${MAX}_$g rc=$?"; grep "box time" "$OUT/main_n${MAX}_$g.txt"
done
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvidia_smi.csv" 2>&1
echo done
