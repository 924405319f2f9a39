#!/bin/bash
# One GPU call for the pipelined light-row kernel (cg_pipe.cu): parity suite under IALS_LIGHT=pipe,
# bit-equality against the cg_rows kernel, then A/B bench runs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
IALS_LIGHT=pipe timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_dist.py -m gpu -x -q > gpurun_out/t_pipe.log 2>&1
echo "rc=$?" >> gpurun_out/t_pipe.log; tail -n 12 gpurun_out/t_pipe.log
for cfg in "A=0" "IALS_LIGHT=pipe" "IALS_LIGHT=pipe IALS_ROWS_PER_WARP=2" "IALS_LIGHT=pipe IALS_ROWS_PER_WARP=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python tools/profile_epoch.py --scale 0.25 --epochs 2 --dump gpurun_out/dump_$tag.npz > gpurun_out/dump_$tag.log 2>&1
  echo "dump [$cfg] rc=$?"
done
python - <<'PY'
import glob, numpy as np
ref = np.load("gpurun_out/dump_A_0.npz")
for f in sorted(glob.glob("gpurun_out/dump_IALS*.npz")):
    d = np.load(f)
    for k in ("user", "item"):
        same = np.array_equal(ref[k], d[k])
        err = np.abs(ref[k] - d[k]).max() / np.abs(ref[k]).max()
        print(f, k, "bit-identical" if same else f"DIFFERENT max rel {err:.3e}")
PY
rm -f gpurun_out/dump_*.npz
tools/gpu_ab.sh "A=0" "IALS_LIGHT=pipe" "IALS_LIGHT=pipe IALS_ROWS_PER_WARP=2" "IALS_LIGHT=pipe IALS_PIPE_SINGLE=100000" \
  "IALS_LIGHT=pipe IALS_HEAVY_THRESHOLD=4096" "IALS_LIGHT=pipe IALS_HEAVY_THRESHOLD=1024" "IALS_LIGHT=pipe IALS_PIPE_SINGLE=128"
