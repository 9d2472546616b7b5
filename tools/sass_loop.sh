#!/bin/bash
# Disassemble the decode kernels of a built library with inline source lines:  tools/sass_loop.sh [lib.so] [outdir]
SO=${1:-lzma-sdk-zliblike_b200/lib/liblzb200.so}
OUT=${2:-/tmp/sass}
mkdir -p $OUT
SOABS=$(readlink -f $SO)
( cd $OUT && rm -f *.cubin && cuobjdump -xelf all $SOABS > /dev/null && nvdisasm -gi lzb_kernels.sm_100a.cubin > all.txt )
grep -n "^\.text\." $OUT/all.txt
