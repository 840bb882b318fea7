#!/bin/bash
# usage: tools/sweep.sh <workload> "<extra args variants separated by ;>"   (GPU box helper, prints one summary line per variant)
wl=$1; shift
IFS=';' read -ra VARS <<< "$1"
for v in "${VARS[@]}"; do
  out=$(timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --no-e2e --no-cpu $v 2>&1 | tail -1)
  python - "$wl" "$v" "$out" <<'PY'
import json,sys
wl,v,out=sys.argv[1:4]
try:
    d=json.loads(out); r=d["roofline"]; p=d["config"]["plan"]
    print(f"{wl} [{v}] fps={d['value']:.0f} ms/step={d['ms_per_step']:.4f} kern_ms={r['kernel_ms']:.4f} GB/s={r['achieved']:.0f} frac={r['frac']:.3f} plan=thr{p['threads']} occ{p['blocks_per_sm']} tile{p['tile_px']} st{p['stages']} k{p.get('kernel',0)} regs{p['regs']} tiles{p['tiles']} seg{p['segments']} clk={d['clocks']['sm_mhz']}")
except Exception as e:
    print(wl, v, "FAILED", out[-300:])
PY
done
