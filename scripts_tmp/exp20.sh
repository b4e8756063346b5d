#!/bin/bash
P="python profiles/tools/small.py"
$P base
RC_TAIL_THRESH=24 $P tail24
RC_TAIL_THRESH=32 $P tail32
RC_SEED_DEPTH=3 $P seed3
$P base
RC_TAIL_THRESH=32 $P tail32
