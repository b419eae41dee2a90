#!/bin/bash
# development iteration: parity sanity check + bench lines of both tensor-core modes (no timeline)
timeout 100 python tools/quick_check.py 2>&1 | grep -E "pred err|Error|error" | tail -8
for m in bf16x3 bf16; do
timeout 100 python bench.py --steps 30 --warmup 3 --mode $m --no-cpu-baseline --no-config4 --train-steps 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', 'timesteps/s %.0f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'K1 us %.1f'%(1e3*d['roofline']['kernel_ms']), 'K2 us %.1f'%(1e3*d['roofline']['mlp_kernel_ms']), 'e2e %.0f'%d['e2e']['value'])"
done
