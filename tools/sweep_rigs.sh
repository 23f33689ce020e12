#!/bin/bash
# throughput and chain latency against the number of concurrent rigs / grow CTAs per image (diagnostic; gpurun -- bash tools/sweep_rigs.sh)
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
for P in 1 2 4 8 13 20; do run p$P A=1; done
P=20; run p20_g10 OLF_LSD_GROW_BLOCKS=10
P=20; run p20_g40 OLF_LSD_GROW_BLOCKS=40
P=30; run p30_g10 OLF_LSD_GROW_BLOCKS=10
