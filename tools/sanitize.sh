#!/bin/bash
# compute-sanitizer passes over the small-n protocol tests (run under gpurun): memcheck for out-of-bounds / misaligned
# accesses, racecheck for shared-memory hazards of the six-lane engine (__syncwarp) the three-warp teams (__syncthreads),
# the lane teams (xt.cuh) and the MSM block trees.
O=gpurun_out
T="tests/test_gpu_protocols.py::test_aggregate_proofs_bytes_match_oracle tests/test_gpu_verify.py::test_verify_aggregate_proof tests/test_gpu_msm.py::test_folds tests/test_gpu_msm.py::test_scalings_match_oracle tests/test_gpu_msm.py::test_msm_signed_digit_corner_scalars tests/test_gpu_msm.py::test_msm_g2_matches_oracle tests/test_gpu_setup.py::test_structured_generators_match_oracle tests/test_gpu_setup.py::test_gipa_checked_prove"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -x -q > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest "tests/test_gpu_protocols.py::test_aggregate_proofs_bytes_match_oracle[2]" tests/test_gpu_msm.py::test_folds "tests/test_gpu_msm.py::test_msm_g1_matches_oracle[300]" -x -q > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log
tail -5 $O/sanitize_memcheck.log; tail -8 $O/sanitize_racecheck.log
