#!/bin/bash
# r02ay: the whole GPU suite one last time (after the device metrics / embeddings paths).
mkdir -p gpurun_out
timeout 160 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 5 gpurun_out/t_all.log
