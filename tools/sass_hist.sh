#!/bin/bash
# tools/sass_hist.sh OBJ [kernel-regex] — opcode histogram per kernel from `cuobjdump -sass` (static instruction mix)
OBJ=${1:?object or .so}
PAT=${2:-.}
cuobjdump -sass "$OBJ" | awk -v pat="$PAT" '
  /Function :/ { fn=$NF; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// { op=$2; sub(/\..*/,"",op); sub(/;$/,"",op); if (fn ~ pat) c[fn"\t"op]++ }
  END { for (k in c) print c[k]"\t"k }' | sort -t$'\t' -k2,2 -k1,1nr
