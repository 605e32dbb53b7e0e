#!/bin/bash
# usage: tools/sass_mix.sh <object> <mangled kernel name> [top-N]  -- SASS opcode histogram of one kernel
cuobjdump -sass -fun "$2" "$1" > /tmp/k.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-16} | tr '\n' ' '
echo
echo "total: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass)"
