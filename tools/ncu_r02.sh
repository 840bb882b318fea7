#!/bin/bash
# Round-2 ncu evidence (GPU box helper; one gpurun call).  Every capture follows a plain run of the same command line.
#   launch list of the default bench | --set full of one clip-kernel launch for the 8-GPU shard shapes | DRAM bytes of one
#   launch at the full BASELINE sizes (a single-pass metric set: no replay, so the 119 GB clips need no save/restore)
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-stream --sustained-s 0"
run_full() {  # tag, bench args...
    tag=$1; shift
    $B --rows "" "$@" > gpurun_out/plain_$tag.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:clip_kernel -s 3 -c 1 -f -o gpurun_out/prof_$tag $B --rows "" "$@" > gpurun_out/ncu_$tag.log 2>&1
    echo "ncu full $tag exit $?"
    # gpurun copies back at most 64 MiB: keep the metric table and the per-instruction source page, not the report itself
    ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
    [ "$tag" = "c4shard" ] || rm -f gpurun_out/prof_$tag.ncu-rep
    ls -la gpurun_out/prof_$tag* | awk '{print $5, $9}'
}
run_dram() {  # tag, bench args...
    tag=$1; shift
    $B --rows "" "$@" > gpurun_out/plain_$tag.log 2>&1 && \
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:clip_kernel -s 3 -c 1 --csv \
        --log-file gpurun_out/dram_$tag.csv $B --rows "" "$@" > gpurun_out/ncu_$tag.log 2>&1
    echo "ncu dram $tag exit $?"
}
$B --rows c3 > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv $B --rows c3 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
run_full c4shard --workload c4 --frames 450
run_full c4pshard --workload c4p --frames 450
run_full c5oshard --workload c5o --frames 150
run_dram c4 --workload c4
run_dram c4p --workload c4p
run_dram c5o --workload c5o
run_dram c5p --workload c5p
run_dram c2 --workload c2
run_dram c3 --workload c3
