#!/bin/bash
# Installs the UNMODIFIED reference (python package + its own CPU library, built by its own CMake through
# scikit-build-core) into baseline/_ref, plus a copy of its test-suite, from a scratch copy of /root/reference
# (which is read-only).  baseline/_ref is git-ignored and travels to the GPU box.  Used as
#   * the conformance suite: the reference's own tests run against our package through shim/ (tools/run_reference_tests.sh)
#   * the loader test: libbitsandbytes_b200.so loaded by the reference's cextension.py (tests/test_gpu_reference_loader.py)
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
if [ -f "$ROOT/baseline/_ref/bitsandbytes/cextension.py" ] && [ -d "$ROOT/baseline/_ref_tests/tests" ]; then
  echo "baseline/_ref already populated"; exit 0
fi
[ -f "$REF/pyproject.toml" ] || { echo "no reference checkout at $REF"; exit 0; }
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/ref"
mkdir -p "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
       --target "$ROOT/baseline/_ref" "$TMP/ref" > "$TMP/install.log" 2>&1 || { tail -20 "$TMP/install.log"; exit 1; }
rm -rf "$ROOT/baseline/_ref_tests"
mkdir -p "$ROOT/baseline/_ref_tests"; cp -r "$REF/tests" "$ROOT/baseline/_ref_tests/tests"
rm -rf "$TMP"
echo "reference installed into baseline/_ref"
