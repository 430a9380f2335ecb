#!/bin/bash
# usage: tools/sass_mix.sh lib.so 'function-substring'   -- static opcode histogram of one kernel
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{on=(index($0,pat)>0)} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\/\*[0-9a-f]{4}\*\///; s/\/\*.*//' | awk '{if($1 ~ /^@/) op=$2; else op=$1; sub(/;/,"",op); n=split(op,a,"."); o=a[1]; if(o=="LDS"||o=="STS"||o=="LDG"||o=="STG") o=a[1]"."a[2]; c[o]++; t++} END{for(k in c) print c[k], k; print t, "TOTAL"}' | sort -rn | head -${3:-30}
