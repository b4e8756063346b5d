#!/bin/bash
P="python profiles/tools/small.py"
$P poll200
RC_LIB=$PWD/ab/librobocheck_p0.so $P poll200_then_1000
RC_LIB=$PWD/ab/librobocheck_p1.so $P poll200_then_2000
RC_LIB=$PWD/ab/librobocheck_p2.so $P poll100
$P poll200
RC_LIB=$PWD/ab/librobocheck_p0.so $P poll200_then_1000
