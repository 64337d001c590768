for C in 4096 2048 1024; do B200Q_HOST_CHUNK=$C python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $C value %.0f e2e %.0f ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']))"; done
python - <<'PY'
import torch,time
x=torch.empty(16384,3,32,32).pin_memory(); y=torch.empty_like(x,device='cuda')
for n in (1,2,4,8):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for c in range(n):
            k=16384//n; y[c*k:(c+1)*k].copy_(x[c*k:(c+1)*k],non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('H2D pinned %d chunks: %.1f GB/s'%(n, 5*x.numel()*4/e0.elapsed_time(e1)/1e6))
PY
