#!/bin/bash
# code size (bytes) of the device functions inside one kernel of a built library: tools/codesize.sh <lib.so> <kernel-name-substring>
set -e
T=$(mktemp -d); cd $T
cuobjdump -xelf all "$1" >/dev/null 2>&1
readelf -sW *.cubin 2>/dev/null | awk '$4=="FUNC"{print $3, $8}' | grep "$2" | sort -n | sed 's/\$_ZN3bls[0-9a-z_]*I[^$]*\$//' | cut -c1-150
rm -rf $T
