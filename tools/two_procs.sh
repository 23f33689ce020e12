#!/bin/bash
# is the ceiling per process (host side) or per GPU?  two bench processes share one GPU, with and without MPS (diagnostic)
one() { tag=$1; P=$2; python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines $P > gpurun_out/two_$tag.log 2>gpurun_out/two_$tag.err; }
show() { python - "$@" <<'PY'
import json,sys
tot=0
for tag in sys.argv[1:]:
    l=[x for x in open(f'gpurun_out/two_{tag}.log') if x.startswith('{')]
    if not l: print(tag,'FAILED'); continue
    d=json.loads(l[-1]); tot+=d['value']; print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'call_p50', d['config']['steadiness']['resident']['call_ms']['p50'])
print('sum', round(tot,1))
PY
}
echo "== no MPS, 2 x 10 rigs"; one a 10 & one b 10 & wait; show a b
which nvidia-cuda-mps-control || exit 0
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log; mkdir -p /tmp/mps /tmp/mps_log
nvidia-cuda-mps-control -d && sleep 2
echo "== MPS, 2 x 10 rigs"; one c 10 & one d 10 & wait; show c d
echo "== MPS, 2 x 20 rigs"; one e 20 & one f 20 & wait; show e f
echo quit | nvidia-cuda-mps-control; sleep 1
