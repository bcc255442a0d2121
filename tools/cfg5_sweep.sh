#!/bin/bash
# BASELINE.json configs[4]: Context(16383,64) multiply+decrypt sweep: T x T operands, the left one
# sharded over the ranks (strong scaling: same total work at every GPU count).  Usage: tools/cfg5_sweep.sh N_GPUS "T list"
# One JSON line per point is appended to gpurun_out/cfg5_sweep_${N}gpu.jsonl.
N=${1:-1}; TS=${2:-"1000 2000 4000"}
OUT=gpurun_out/cfg5_sweep_${N}gpu.jsonl; : > $OUT
for T in $TS; do
  PAIRS=2; if [ $((T * T / N)) -ge 16000000 ]; then PAIRS=1; fi
  if [ "$N" = "1" ]; then
    python bench.py --workload cfg5 --t1 $T --t2 $T --pairs $PAIRS --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' >> $OUT
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + T / 100)) \
      bench.py --gpus $N --workload cfg5 --t1 $T --t2 $T --pairs $PAIRS --scaling strong --steps 5 --warmup 3 2>/dev/null | grep '^{' >> $OUT
  fi
done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print("N=%d %s: %.4g blocks/s, %.3f ms/step, mul %.0f GB/s (%.3f), dec %.0f GB/s (%.3f), e2e %.4g" % (
        d["n_gpus"], d["config"]["workload"][6:40], d["value"], d["ms_per_step"], d["kernels"]["multiply"]["gbs"],
        d["kernels"]["multiply"]["frac_of_peak"], d["kernels"]["decrypt"]["gbs"], d["kernels"]["decrypt"]["frac_of_peak"], d["e2e"]["value"]))
PY
