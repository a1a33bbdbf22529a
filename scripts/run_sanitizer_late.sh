#!/bin/bash
# compute-sanitizer over the kernels added after the full pass of scripts/run_sanitizer.sh: the stride-1 Conv4d tile kernel, the
# layer-9 GEMM with query_embed in its producer warps, the compact hidden image (written by GEMM1, value plane derived by layer 10),
# the compact gather form, the sub-tile pipelined and weight-stationary experiment kernels, the tcgen05 correlation.
mkdir -p gpurun_out
SEL='conv4d_block or operators_match_torch or key_and_round2 or (matches_reference_golden and render_64) or compact or operand_image_chain or linear_tc'
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_render_gpu.py tests/test_gather_gpu.py tests/test_ufc_gpu.py tests/test_ufc_native_gpu.py -x -q -m gpu -k "$SEL and not cta-pairs and not cluster" -p no:cacheprovider > gpurun_out/r2_sanitizer_late_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2_sanitizer_late_memcheck.log | tail -5
RSEL='key_and_round2 or (test_matches_reference_golden and render_64_frontal) or compact_hidden'
timeout 700 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k "$RSEL" -p no:cacheprovider > gpurun_out/r2_sanitizer_late_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_late_racecheck.log | tail -5
