#!/bin/bash
# compute-sanitizer over the kernels that changed in round 2 (run under gpurun; small cases)
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 9"
run() {  # name, tool, pytest selection
  echo "== $1 ($2): $3 -k \"$4\"" >> gpurun_out/r2_sanitizer.log
  timeout 900 $S --tool $2 python -m pytest $3 -k "$4" -m gpu -q -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" | tail -6 >> gpurun_out/r2_sanitizer.log
}
: > gpurun_out/r2_sanitizer.log
run free-running memcheck tests/test_gpu_ragged.py "(prefixes and 11) or iteration_cap or warmup_adapts or streamed"
run free-running racecheck tests/test_gpu_ragged.py "(prefixes and 11) or iteration_cap"
run one-shot-ragged memcheck tests/test_gpu_ragged.py "one_shot"
run streaming memcheck tests/test_gpu_streaming.py "streamed_summaries_equal or one_shot_summary"
run streaming racecheck tests/test_gpu_streaming.py "streamed_summaries_equal"
run device-source memcheck tests/test_gpu_device_source.py "(bitwise and 10) or (orbits and 7) or not_built_in"
run logistic memcheck tests/test_gpu_parity.py "logistic_gradient_operator"
run chain-kernel racecheck tests/test_gpu_parity.py "seeded_trajectories"
cat gpurun_out/r2_sanitizer.log
