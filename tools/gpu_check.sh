#!/bin/bash
# Sequential GPU checks with short timeouts; every log lands in gpurun_out/ even on a hang.
#   tools/gpu_check.sh [tests] [bench] [launches] [ncu]
mkdir -p gpurun_out
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" >> "gpurun_out/$name.log"
  echo "== $name rc=$rc"; tail -n "${TAIL:-6}" "gpurun_out/$name.log"
  return $rc
}
what=${*:-tests bench launches ncu}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for w in $what; do
  case $w in
    tests) run t_all 900 python -m pytest tests -m gpu -x -q || exit 1 ;;
    bench) run bench 600 python bench.py --steps 10 --warmup 3 --cpu-epochs 2
      python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench.log") if l.startswith("{")][-1])
r=d["roofline"]
print("ms/epoch", d["ms_per_step"], "frac(light kernel)", r["frac"], "whole-solve frac", r["whole_solve"]["frac"])
print({k: round(v, 3) for k, v in r["phases_ms_per_epoch"].items()})
print("cpu", d["cpu_baseline"]["ms_per_epoch"], "cores", d["cpu_baseline"]["cores"], "e2e ms", d["e2e"]["ms_per_step"], d["clocks"])
PY
      ;;
    benchN) N=${NGPU:-2}; run bench$N 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ;;
    launches) TAIL=3 run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/launches.csv python tools/profile_epoch.py --epochs 3 --recommend 4096 ;;
    ncu) TAIL=3 run ncu_rows 900 ncu --set full --clock-control none --import-source on -k regex:cg_rows_kernel -s 2 -c 2 \
        -f -o gpurun_out/prof_rows python tools/profile_epoch.py --epochs 2
      TAIL=3 run ncu_wgram 900 ncu --set full --clock-control none --import-source on -k regex:wgram_kernel -s 5 -c 1 \
        -f -o gpurun_out/prof_wgram python tools/profile_epoch.py --epochs 2 ;;
  esac
done
