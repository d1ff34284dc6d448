#!/bin/bash
# developer aid: rebuild, dump the SASS of the default state kernel to $1 and print its loop structure
set -e
cd /root/repo
python __graft_entry__.py 2>&1 | tail -1
P=bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200
K=${2:-_Z17csmc_state_kernelILi2ELi1ELb0ELi256ELi2ELi2ELb0EEv9StateArgs}
cuobjdump -res-usage $P/build/sweep.o 2>/dev/null | grep -A1 "$K" | tail -1
cuobjdump -sass -fun "$K" $P/build/sweep.o > $1
python3 /root/repo/tools/sass_loops.py $1
