#!/usr/bin/env bash
# Install the UNMODIFIED reference (yakvrz/minesweeper-ppo) into baseline/_ref/.
#
# The reference has no setup.py / pyproject.toml (its scripts run with PYTHONPATH=., README.md:28-32),
# so "install" = a verbatim copy of the tree.  baseline/_ref/ is git-ignored (never part of this
# repo's history) but NOT gpurun-ignored, so it travels to the GPU box with the working tree, where
#   * tests/test_gpu_reference_live.py drives it lock-step against the CUDA env, and
#   * bench.py times its numba env on the box's host cores (cpu_baseline.reference_numba).
# __graft_entry__.build() runs this whenever /root/reference is present.
set -euo pipefail
SRC="${MSW_REFERENCE_SRC:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/minesweeper" ]; then
  echo "install_reference: $SRC not found (nothing to install)" >&2
  exit 3
fi
rm -rf "$DST.tmp"
mkdir -p "$DST.tmp"
# sources only: no caches, no web UI assets, no docs
( cd "$SRC" && find . -type f \( -name '*.py' -o -name '*.yaml' -o -name 'requirements.txt' \) \
    -not -path './webui/*' -not -path '*/__pycache__/*' -print0 | cpio -0 -pdm --quiet "$DST.tmp" ) 2>/dev/null \
  || ( cd "$SRC" && find . -type f \( -name '*.py' -o -name '*.yaml' -o -name 'requirements.txt' \) \
    -not -path './webui/*' -not -path '*/__pycache__/*' | while read -r f; do mkdir -p "$DST.tmp/$(dirname "$f")"; cp "$f" "$DST.tmp/$f"; done )
( cd "$SRC" && find . -type f -name '*.py' -not -path './webui/*' -not -path '*/__pycache__/*' | sort | xargs sha256sum ) > "$DST.tmp/SHA256SUMS"
rm -rf "$DST"
mv "$DST.tmp" "$DST"
echo "installed reference into $DST ($(find "$DST" -name '*.py' | wc -l) python files)"
