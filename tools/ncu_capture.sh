#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <bench args...>  : plain run, then ncu --set full on one clip_kernel launch (GPU box helper)
tag=$1; shift
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:clip_kernel -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/ncu_$tag.log 2>&1
echo "ncu $tag exit $?"
