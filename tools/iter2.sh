for m in bf16x3; do
timeout 100 python bench.py --steps 30 --warmup 3 --mode $m --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', 'timesteps/s %.0f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'K1 us %.1f'%(1e3*d['roofline']['kernel_ms']), 'K2 us %.1f'%(1e3*d['roofline']['mlp_kernel_ms']), 'e2e %.0f'%d['e2e']['value'])"
done
