#!/bin/bash
# compute-sanitizer over the GPU suite (memcheck) and over the tensor-core / render subset (racecheck + synccheck).
# The full-resolution cases (65 536 rays, 512x512, 4096^2 operators) are left out: under the sanitizer every kernel runs 20-100x slower.
#   bash scripts/run_sanitizer.sh            -> gpurun_out/r2_sanitizer_*.log, summary on stdout
mkdir -p gpurun_out
SEL='not full_resolution and not full_image and not config4 and not 512 and not two_gpu and not graph_replay and not batches and not 4096 and not pose_feature_operators'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2_sanitizer_memcheck.log | tail -5
# the opt-in experiment kernels (cta_group::2 pairs, cluster multicast) are left out: racecheck flags the cross-CTA shared-memory
# write of tcgen05.alloc.cta_group::2 itself; the default path is what ships
RSEL='((matches_reference_golden and render_64) or gemm_tc or gather or key_and_round2 or linear_tc or chunk_invariance) and not cta-pairs and not cluster'
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 python -m pytest tests/test_render_gpu.py tests/test_gather_gpu.py tests/test_ufc_native_gpu.py -x -q -m gpu -k "$RSEL" -p no:cacheprovider > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_racecheck.log | tail -5
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k "(matches_reference_golden and render_64) or gemm_tc" -p no:cacheprovider > gpurun_out/r2_sanitizer_synccheck.log 2>&1
echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_synccheck.log | tail -5
