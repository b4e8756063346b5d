#!/bin/bash
P="python profiles/tools/small.py"
for L in 8 4 12 16 24 32; do RC_SPLIT_LOAD=$L $P load$L; done
