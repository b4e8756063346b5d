#!/bin/bash
echo "== old"; timeout 200 python profiles/tools/brute.py 2>&1 | head -1
export RC_LIB=$PWD/ab/librobocheck_bq.so
echo "== queued"; timeout 200 python profiles/tools/brute.py 2>&1 | head -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "brute or hit_pair or audit or coplanar or narrow_phase or stage_pipeline or tiny_meshes or golden or known_answer or megakernel or scheduling_variants" 2>&1 | tail -3
unset RC_LIB
echo "== old"; timeout 200 python profiles/tools/brute.py 2>&1 | head -1
