#!/bin/bash
# BASELINE.json configs[4]: Context(16383,64) multiply+decrypt sweep: T x T operands, the left one
# sharded over the ranks (strong scaling: same total work at every GPU count).  Usage: tools/cfg5_sweep.sh N_GPUS "T list"
# Two JSON lines per point (fused and two-pass step) are appended to gpurun_out/cfg5_sweep_${N}gpu.jsonl; 50 timed
# steps per point, per-launch medians and minima next to the means.
N=${1:-1}; TS=${2:-"1000 2000 4000"}; STEPS=${STEPS:-50}
OUT=gpurun_out/cfg5_sweep_${N}gpu.jsonl; : > $OUT
for T in $TS; do
  PAIRS=2; if [ $((T * T / N)) -ge 8000000 ]; then PAIRS=1; fi
  for MODE in ${MODES:-fused two-pass}; do
    if [ "$N" = "1" ]; then
      python bench.py --workload cfg5 --t1 $T --t2 $T --pairs $PAIRS --steps $STEPS --warmup 3 --mode $MODE --no-extras --no-cpu-baseline 2>/dev/null | grep '^{' >> $OUT
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + T / 100)) \
        bench.py --gpus $N --workload cfg5 --t1 $T --t2 $T --pairs $PAIRS --scaling strong --steps $STEPS --warmup 3 --mode $MODE --no-extras 2>/dev/null | grep '^{' >> $OUT
    fi
  done
done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    k = d["kernels"]
    parts = []
    for name in ("fused_multiply_decrypt", "multiply", "decrypt"):
        if name in k:
            parts.append("%s %.3f of peak (us/launch: mean %.1f median %.1f min %.1f)" % (
                name, k[name]["frac_of_peak"], k[name]["avg_launch_us"], k[name]["median_step_launch_us"], k[name]["min_step_launch_us"]))
    print("N=%d %s [%s]: %.4g blocks/s, %.3f ms/step, e2e %.4g | %s" % (
        d["n_gpus"], d["config"]["workload"][6:46], d["config"]["mode"][:8], d["value"], d["ms_per_step"], d["e2e"]["value"], "; ".join(parts)))
PY
