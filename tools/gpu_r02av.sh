#!/bin/bash
# r02av: ials_trainer_recommend_embeddings (cold users scored against the resident item factors): parity.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_score_tc.py tests/test_id_mapping.py tests/test_evaluator_wide.py tests/test_gpu_parity.py -m gpu -q -x \
  -k "not full_size and not c1_config and not c3_" > gpurun_out/t_av.log 2>&1
echo "== tests rc=$?"; tail -n 6 gpurun_out/t_av.log
