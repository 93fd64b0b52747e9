#!/bin/bash
# usage: sass_mix.sh <object-or-so> <mangled-name-substring>   -> instruction mix of one kernel's SASS
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ { on = index($0, pat) > 0; next }
on && /^ +\/\*[0-9a-f]+\*\/ / { ins=$2; if (ins ~ /^@/) ins=$3; sub(/\..*/, "", ins); sub(/;$/, "", ins); c[ins]++; n++ }
END { printf("total %d\n", n); for (k in c) printf("%6d %s\n", c[k], k) | "sort -rn" }'
