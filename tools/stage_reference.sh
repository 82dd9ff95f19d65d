#!/bin/sh
# Stages the reference's pure-Python package files (no compiled modules) into the git-ignored
# scratch/refstage/ so that tests/test_overlay.py can run the UNMODIFIED reference front end on
# top of librhfq_b200.so on the GPU box, which has no /root/reference.  Nothing staged here is
# ever committed (scratch/ is in .gitignore).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$ROOT/scratch/refstage"
rm -rf "$DST"
mkdir -p "$DST"
(cd "$SRC" && find reheatfunq -name '*.py' | while read f; do mkdir -p "$DST/$(dirname "$f")"; cp "$f" "$DST/$f"; done)
echo "staged $(find "$DST" -name '*.py' | wc -l) files into $DST"
