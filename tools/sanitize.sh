#!/bin/bash
# compute-sanitizer over the GPU parity tests of the four kernel families (eigensum pair kernels, homogeneous spaces, Gram, Cholesky),
# one tool per gpurun call (B200_PROFILING.md):   gpurun -- 'bash tools/sanitize.sh memcheck|racecheck|synccheck'
# Full-size tests are excluded (the tools slow kernels down 10-100x); logs go to gpurun_out/sanitizer_<tool>.log.
TOOL=${1:-memcheck}
SEL='golden or ragged or layouts or gram_matches or rank_k or column_block or empty or cholesky_matches or right_hand_sides or leading_dimension or sharded or stiefel53 or fused_stiefel or so5_quotients'
FILES="tests/test_eigensum_gpu.py tests/test_homogeneous_gpu.py tests/test_random_phase_gpu.py tests/test_gpr_gpu.py tests/test_noncompact_gpu.py"
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool "$TOOL" --error-exitcode 9 --launch-timeout 0 \
    python -m pytest $FILES -x -q -m gpu -k "$SEL" > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_$TOOL.log
grep -E "ERROR SUMMARY|passed|failed|exit code|Race reported|hazard" gpurun_out/sanitizer_$TOOL.log | tail -8
