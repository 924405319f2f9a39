#!/bin/bash
# r02aw: the whole GPU suite and smoke() once more after the last changes (embeddings path, two new tests).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 5 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
