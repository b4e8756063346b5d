#!/bin/bash
# A/B of several builds of the library on one box: ab_libs.sh lib_a.so lib_b.so ...  (two interleaved repetitions)
for rep in 1 2; do for lib in "$@"; do RC_LIB=$PWD/$lib python profiles/tools/quick.py $(basename $lib .so) $QUICK_ARGS 2>&1 | grep -v "^$"; done; done
