#!/bin/bash
# The reference's OWN test-suite (baseline/_ref_tests, copied by tools/install_reference.sh) against this package
# through the import shim (shim/bitsandbytes -> bitsandbytes_b200), on the GPU:  SURVEY.md appendix B.
#   usage (GPU box): bash tools/run_reference_tests.sh [extra pytest args]
# Writes gpurun_out/reftests_<file>.log and a summary table gpurun_out/reftests_summary.txt.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/gpurun_out"
cd "$ROOT/baseline/_ref_tests" || { echo "baseline/_ref_tests missing (tools/install_reference.sh)"; exit 0; }
export PYTHONPATH="$ROOT/shim:$ROOT:$ROOT/tools/ref_test_stubs" BNB_TEST_DEVICE=cuda
: > "$ROOT/gpurun_out/reftests_summary.txt"
for f in ${REFTESTS:-test_ops test_functional test_linear4bit test_linear8bitlt test_autograd test_modules test_parametrize test_optim}; do
  timeout 1500 python -m pytest "tests/$f.py" -q -p no:cacheprovider --maxfail=10000 -k "not benchmark" "$@" \
      > "$ROOT/gpurun_out/reftests_$f.log" 2>&1
  echo "$f: $(tail -1 "$ROOT/gpurun_out/reftests_$f.log")" | tee -a "$ROOT/gpurun_out/reftests_summary.txt"
done
