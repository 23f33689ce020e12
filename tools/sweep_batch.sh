#!/bin/bash
# frames per call against rigs (diagnostic)
run() { tag=$1; shift; env $ENVV python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline "$@" > gpurun_out/sweep_$tag.log 2>gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/sweep_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'call_p50', c['steadiness']['resident']['call_ms']['p50'], 'rig_call_ms', c['rig_call_ms']['total'], 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
}
ENVV="A=1"; run p10_b8 --pipelines 10 --batch 8
ENVV="A=1"; run p20_b8 --pipelines 20 --batch 8
ENVV="A=1"; run p26_b2 --pipelines 26 --batch 2
ENVV="OLF_LSD_GRAPH=0"; run p20_nograph --pipelines 20
