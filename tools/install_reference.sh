#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with the
# repo snapshot): the package by pip (--target), its experiments/ scripts and tests/ by plain copy next
# to it.  Used by bench.py --impl reference / cpu_baseline (kind "reference") and by
# tests/test_reference_scripts.py (the reference's own scripts and tests run against the sleekit/ shim).
# Nothing from the reference enters the repository's history.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
rm -rf /tmp/_slk_refcopy && cp -r "$REF" /tmp/_slk_refcopy
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" /tmp/_slk_refcopy
cp -r "$REF/experiments" "$ROOT/baseline/_ref/experiments"
cp -r "$REF/tests" "$ROOT/baseline/_ref/reference_tests"
rm -rf /tmp/_slk_refcopy
echo "installed into $ROOT/baseline/_ref"
