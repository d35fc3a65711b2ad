#!/usr/bin/env bash
# Places a VERBATIM, unmodified copy of the reference's Python package and scripts under oracle/_ref/ so that the
# reference itself can run on the GPU box (gpurun ships /root/repo, not /root/reference): it is the CPU baseline
# `bench.py --impl reference` times (cpu_baseline.kind = "reference") and the class set the drop-in GPU tests drive
# through rpo_b200.algo.install().  oracle/_ref/ is git-ignored (never part of the history) but not gpurun-ignored.
# TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under rpo_b200/ imports it.
#   usage: oracle/make_ref_copy.sh [reference_root=/root/reference]
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
[ -f "$SRC/rpo/env/electrical_grid/evopf.py" ] || { echo "no reference tree at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/rpo" "$SRC/scripts" "$DST/"
[ -f "$SRC/LICENSE" ] && cp "$SRC/LICENSE" "$DST/"
chmod -R u+w "$DST"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
( cd "$SRC" && find rpo scripts -type f ! -path "*/__pycache__/*" -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "reference copied to $DST ($(du -sh "$DST" | cut -f1)); $(wc -l < "$DST/SHA256SUMS") files"
