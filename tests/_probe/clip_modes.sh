#!/bin/bash
for m in 0 1 2 3 7 16 17 19; do echo "== CRT_CLIP_RELEASE=$m"; CRT_CLIP_RELEASE=$m python tests/_probe/clip_diff.py ${1:-default} 30 2>&1 | grep -E "vs-clip" | cut -c1-150; done
