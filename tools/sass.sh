#!/bin/bash
# usage: tools/sass.sh <mangled-substring>   -> compact SASS listing of the first matching kernel in libpixsht.so
cuobjdump -sass "$(dirname "$0")/../pixell.jl_b200/lib/libpixsht.so" | awk -v pat="$1" '/Function :/{f=(index($0,pat)>0)} f' \
  | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#^\s+/\*([0-9a-f]{4})\*/\s+#\1 #; s#\s*/\*.*##'
