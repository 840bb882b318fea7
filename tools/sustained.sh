#!/bin/bash
# usage: tools/sustained.sh <workload> <steps> "<args variants separated by ;>"
wl=$1; steps=$2; shift 2
IFS=';' read -ra VARS <<< "$1"
for v in "${VARS[@]}"; do
  timeout 300 python bench.py --workload $wl --steps $steps --warmup 5 --no-e2e --no-cpu $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; p=d['config']['plan']
print('$wl steps=$steps [$v]', 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kern', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'probe', round(r['stream_probe_GBps']), 'k%d st%d thr%d regs%d' % (p['kernel'], p['stages'], p['threads'], p['regs']), 'clk', d['clocks']['sm_mhz'], 'W', round(d['clocks']['power_w_max'] or 0))"
done
