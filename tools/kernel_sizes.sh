#!/bin/bash
# Code size (SASS bytes), registers and stack of every kernel in libreseq_b200.so:  tools/kernel_sizes.sh [pattern]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP=$(mktemp -d); cd "$TMP"
cuobjdump -xelf all "$ROOT/reseq_b200/libreseq_b200.so" > /dev/null
readelf -S -W *.cubin 2>/dev/null | grep " \.text\." | python3 -c "
import sys
rows=[]
for l in sys.stdin:
    t=l.split(); i=[k for k,x in enumerate(t) if x.startswith('.text.')][0]
    rows.append((int(t[i+4],16), t[i][6:]))
for s,n in sorted(rows): print('%8.1f KB  %s' % (s/1024, n[:110]))
" | grep -E "${1:-.}"
cuobjdump --dump-resource-usage "$ROOT/reseq_b200/libreseq_b200.so" 2>/dev/null | grep -A1 -E "Function" | grep -E "Function|REG" | paste - - | sed -E 's/.*Function ([^:]*):.*REG:([0-9]+) STACK:([0-9]+) SHARED:([0-9]+).*/REG \2 STACK \3 SHARED \4  \1/' | cut -c1-150 | grep -E "${1:-.}"
rm -rf "$TMP"
