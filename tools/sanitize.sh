#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small parity subset (run on the GPU box); logs -> gpurun_out/
# usage: tools/sanitize.sh [tag]
tag=${1:-r2}
mkdir -p gpurun_out
SEL="test_known_answers or test_edge_cases or test_gated_batch_of_tiny_ragged_windows or (test_config_parity and live) or test_invalid_view_is_fail_safe"
SEL2="test_device_bound_known_answers_and_batches or test_compaction_matches_the_reference_semantics or test_mirror_drives_the_compaction_after_a_window or test_mirror_follows_deltas_and_applies_the_deletion"
for tool in memcheck racecheck; do
  MSS_WATCHDOG_MS=600000 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --log-file gpurun_out/${tag}_sanitizer_${tool}.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/${tag}_sanitizer_${tool}.pytest.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/${tag}_sanitizer_${tool}.pytest.log
  tail -3 gpurun_out/${tag}_sanitizer_${tool}.log
  MSS_WATCHDOG_MS=600000 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --log-file gpurun_out/${tag}_sanitizer_${tool}_2.log \
      python -m pytest tests/test_dual_bound.py tests/test_compact.py tests/test_mirror.py -m gpu -x -q -k "$SEL2" > gpurun_out/${tag}_sanitizer_${tool}_2.pytest.log 2>&1
  echo "$tool (bound / compact / mirror) exit $?" | tee -a gpurun_out/${tag}_sanitizer_${tool}_2.pytest.log
  tail -3 gpurun_out/${tag}_sanitizer_${tool}_2.log
done
