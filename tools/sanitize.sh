#!/bin/bash
# compute-sanitizer passes over the small-n protocol tests (run under gpurun): memcheck for out-of-bounds / misaligned
# accesses, racecheck for shared-memory hazards of the six-lane engine (__syncwarp) and the three-warp teams (__syncthreads).
O=gpurun_out
T="tests/test_gpu_protocols.py::test_aggregate_proofs_bytes_match_oracle tests/test_gpu_verify.py::test_verify_aggregate_proof tests/test_gpu_msm.py::test_folds"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -x -q > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest "tests/test_gpu_protocols.py::test_aggregate_proofs_bytes_match_oracle[2]" tests/test_gpu_msm.py::test_folds -x -q > $O/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitize_racecheck.log
tail -5 $O/sanitize_memcheck.log; tail -8 $O/sanitize_racecheck.log
