#!/bin/bash
# compute-sanitizer passes over the small GPU tests (golden fixtures, scan incl. the fast path and the statement golden,
# random-stage tests of the support filter / post-processing chain, the rolling submit entry points).
#   bash tools/sanitize.sh > profiles/rNN_sanitizer.txt 2>&1        (on a B200; ~6 min)
SEL='test_cuda_matches_golden or statement_by_statement or test_scan_batch_and_empty_map or test_reprojection or test_support_filter_random or test_rolling_device or (test_postprocess_random and not subsampled and mode0) or general_q'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool  python -m pytest tests -m gpu -q -k \"$SEL\""
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 0 python -m pytest tests -m gpu -q -x -k "$SEL" > /tmp/san_$tool.log 2>&1
  echo "exit code: $?"
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Error:" /tmp/san_$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
  grep -E " and (Read|Write) access" /tmp/san_$tool.log | sed 's/+0x[0-9a-f]*//; s/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head -12
done
for tool in synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool  python -m pytest tests -m gpu -q -k \"test_cuda_matches_golden or test_rolling_device\""
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "test_cuda_matches_golden or test_rolling_device" 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|Uninitialized|Barrier" | head -20
  echo "exit code: ${PIPESTATUS[0]}"
done
