#!/bin/bash
# Sequential GPU checks with short timeouts; every log lands in gpurun_out/ even on a hang.
mkdir -p gpurun_out
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" >> "gpurun_out/$name.log"
  echo "== $name rc=$rc"; tail -n "${TAIL:-6}" "gpurun_out/$name.log"
  return $rc
}
run t_staged 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged or half_steps or empty" || exit 1
run t_all 600 python -m pytest tests -m gpu -x -q || exit 1
run bench 400 python bench.py --steps 10 --warmup 3 || exit 1
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench.log") if l.startswith("{")][-1])
print("ms/epoch", d["ms_per_step"], "solve u/i", d["roofline"]["solve_users_ms"], d["roofline"]["solve_items_ms"], "gram", d["roofline"]["gram_ms_per_epoch"], "frac", d["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["ms_per_epoch"], "cores", d["cpu_baseline"]["cores"], "e2e ms", d["e2e"]["ms_per_step"], d["clocks"])
PY
