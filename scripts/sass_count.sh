#!/bin/bash
# usage: sass_count.sh <object.o> <mangled-kernel-substring>  -> instruction count and opcode histogram of one kernel
OBJ=$1; PAT=$2
cuobjdump -sass "$OBJ" 2>/dev/null | awk -v pat="$PAT" '
  /Function :/ { on = index($0, pat) > 0 }
  on && /^[ \t]+\/\*[0-9a-f]{4}\*\// { sub(/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+/, ""); sub(/[ \t]*\/\*.*$/, ""); print }' > /tmp/sass_one.txt
echo "instructions: $(wc -l < /tmp/sass_one.txt)"
awk '{ op=$1; if (op ~ /^@/) op=$2; sub(/\..*/, "", op); print op }' /tmp/sass_one.txt | sort | uniq -c | sort -rn | head -${3:-12}
