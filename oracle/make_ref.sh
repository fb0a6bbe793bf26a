#!/usr/bin/env bash
# Test infrastructure: installs the UNMODIFIED reference (tulip-control/polytope, pure
# Python) from /root/reference into the git-ignored oracle/_ref/, so that it travels to
# the GPU box with the snapshot (gpurun ships git-ignored files; /root/reference itself
# does not exist there).  Nothing under oracle/_ref/ is ever committed or edited.
#
#   oracle/_ref/polytope/            the installed package (pip --target, wheel built from
#                                    a scratch copy because the build writes version files)
#   oracle/_ref/reference_tests/     the reference's own tests/, run by tests/test_gpu_dropin.py
#                                    against the package with the GPU lpsolve patched in
#
# Used by: tests/ (oracle pinning, drop-in tests), bench.py's cpu_baseline / --impl reference
# legs.  The product package polytope_b200/ never imports it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$SRC/polytope" ]; then
    echo "make_ref.sh: $SRC not present (GPU box?): keeping the prebuilt $OUT" >&2
    exit 0
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$SRC" "$TMP/src"
rm -rf "$OUT"
mkdir -p "$OUT"
python -m pip install --quiet --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$OUT" "$TMP/src" >"$TMP/pip.log" 2>&1 || { cat "$TMP/pip.log" >&2; exit 1; }
mkdir -p "$OUT/reference_tests"
cp "$SRC"/tests/*.py "$SRC"/tests/pytest.ini "$OUT/reference_tests/"
( cd "$SRC" && find polytope tests -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$OUT/SOURCES.sha256"
echo "reference installed into $OUT"
