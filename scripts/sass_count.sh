#!/bin/bash
# static SASS instruction count per kernel of a built library (a proxy for ALU-bound kernels when no GPU is at hand)
/usr/local/cuda/bin/cuobjdump -sass "${1:-chessrl_b200/libchessrl_b200.so}" 2>/dev/null | awk '
  /Function :/ {name=$3}
  /^ +\/\*[0-9a-f]+\*\/ +[@A-Z]/ {cnt[name]++}
  END {for (n in cnt) print cnt[n], n}' | sort -k2
