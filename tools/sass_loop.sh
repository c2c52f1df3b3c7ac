#!/bin/bash
# usage: tools/sass_loop.sh <obj> <mangled-function-substring>   -- dump the cleaned SASS of one kernel to stdout
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" | grep "Function :" | grep "$pat" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$fn" "$obj" | grep ";" | sed -E 's/^\s+//; s/\s+\/\* 0x[0-9a-f]+ \*\/$//; s/\s+/ /g'
