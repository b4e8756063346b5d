#!/bin/bash
# knob sensitivity of the planner-sized batches (us per call)
P="python profiles/tools/small.py"
$P base
RC_SPLIT_MAX=1 $P nosplit
RC_SHARE=0 $P noshare
RC_SPLIT_MAX=1 RC_SHARE=0 $P nosplit_noshare
RC_SPLIT_LOAD=2 $P load2
RC_SPLIT_LOAD=4 $P load4
RC_SPLIT_LOAD=16 $P load16
RC_SPLIT_LOAD=32 $P load32
RC_DONATE_MIN=40 RC_MAIL_SLOTS=8 $P slots8
RC_MAIL_SLOTS=16 $P slots16
RC_CULLED_WARPS=16 $P warps16
RC_CULLED_WARPS=24 $P warps24
RC_PDL=0 $P nopdl
RC_TAIL_THRESH=8 $P tail8
