#!/bin/bash
P="python profiles/tools/small.py"
for W in 32 24 16 32 24 16; do RC_CULLED_WARPS=$W timeout 200 $P warps$W; done
