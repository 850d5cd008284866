#!/bin/bash
# usage: tools/sass_hist.sh <object> <mangled kernel name>: static SASS opcode histogram + ptxas resource line
cuobjdump -sass -fun "$2" "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]{4}\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/[.;].*//' | sort | uniq -c | sort -rn | head -${3:-14}
echo "total: $(cuobjdump -sass -fun "$2" "$1" | grep -cE '^\s+/\*[0-9a-f]{4}\*/')"
