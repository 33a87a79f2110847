#!/bin/sh
# Stage the UNMODIFIED reference implementation of the hot path (pure Python: torchcde, torchdiffeq and the vector field
# of src/ncde) into oracle/_ref/ so that it travels to the GPU box with the snapshot.  oracle/_ref/ is git-ignored (the
# reference's sources never enter this repository's history) but not gpurun-ignored.  Test infrastructure only: the
# directory is imported by bench.py's reference arm / cpu_baseline leg and by nothing in the product package.
#   usage: oracle/make_ref.sh [/root/reference]
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
[ -d "$REF/modules/torchcde/torchcde" ] || { echo "reference not found at $REF; nothing staged"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST/ncde_ref_vector_fields"
cp -r "$REF/modules/torchcde/torchcde" "$DST/torchcde"
cp -r "$REF/modules/torchdiffeq/torchdiffeq" "$DST/torchdiffeq"
# src/ncde/__init__.py imports the un-vendored `autots`; the vector-field base module has no external dependency
cp "$REF/src/ncde/vector_fields/base.py" "$DST/ncde_ref_vector_fields/base.py"
: > "$DST/ncde_ref_vector_fields/__init__.py"
find "$DST" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/REFERENCE_COMMIT"
echo "staged reference hot path into $DST"
