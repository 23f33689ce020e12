#!/bin/bash
# is the host the limiter?  tracker threads, sync mode (diagnostic)
run() { tag=$1; shift; env $ENVV python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines 20 "$@" > gpurun_out/sweep_$tag.log 2>/dev/null
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/sweep_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'rig_call_ms', c['rig_call_ms']['total'], 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2), 'cores', c['host_cores'])
PY
}
ENVV="A=1"; run notrack --no-track
ENVV="A=1"; run track4 --trackers 4
ENVV="OLF_SYNC=block"; run block
ENVV="OLF_SYNC=spin"; run spin
nproc; cat /proc/cpuinfo | grep "model name" | head -1
