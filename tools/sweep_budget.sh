#!/bin/bash
# throughput against the grow-pass step budget (regions parked and re-packed densely every `budget` queue entries); diagnostic
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines $P > gpurun_out/sweep_$tag.log 2>/dev/null
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/sweep_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'rig_call_ms', c['rig_call_ms']['total'], 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
}
P=20; run b0 A=1
for b in 128 256 512 1024; do P=20; run b$b OLF_LSD_GROW_BUDGET=$b; done
P=30; run p30_b256 OLF_LSD_GROW_BUDGET=256
P=30; run p30_b512_g40 OLF_LSD_GROW_BUDGET=512 OLF_LSD_GROW_BLOCKS=40
P=8; run p8_b256 OLF_LSD_GROW_BUDGET=256
