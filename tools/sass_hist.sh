#!/bin/bash
# usage: tools/sass_hist.sh <function-substring> [next-function-substring]  -> SASS opcode histogram
LIB=/root/repo/monte-carlo-project-cuda_b200/libmcb200.so
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function :/{f = index($0, pat) > 0} f{print}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{ if ($2 ~ /^@/) print $3; else print $2 }' | sed 's/;//' | sort | uniq -c | sort -rn
