#!/bin/bash
for f in 0 1 0 1; do RC_FEED=$f timeout 200 python profiles/tools/host_small.py; done
