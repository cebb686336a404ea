#!/bin/bash
# usage: tools/sass_of.sh <object-or-so> <kernel-name-substring>  -> SASS of the matching function, one instruction per line
cuobjdump -sass "$1" | awk -v k="$2" '/Function : /{f = index($0, k) > 0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\* 0x[0-9a-f]* \*\///'
