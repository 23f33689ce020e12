#!/bin/bash
# bench.py at 20 rigs under different environment settings: usage sweep_env.sh tag=VAR=VAL[,VAR=VAL] ...
for spec in "$@"; do
  tag=${spec%%=*}; envs=${spec#*=}; envs=${envs//,/ }
  env $envs python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines ${P:-20} > gpurun_out/sweep_$tag.log 2>gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/sweep_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'call_p50', c['steadiness']['resident']['call_ms']['p50'], 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
done
