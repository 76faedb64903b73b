import torch
dev='cuda:0'
x=torch.empty(1<<30, dtype=torch.bfloat16, device=dev).normal_()  # 2 GiB
y=torch.empty_like(x)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
nb=x.numel()*2
t=timed(lambda: y.copy_(x)); print(f"copy: {2*nb/t/1e6:.0f} GB/s (read+write)")
xi=x.view(torch.int32)
t=timed(lambda: xi.sum()); print(f"sum int32 (read only): {nb/t/1e6:.0f} GB/s")
t=timed(lambda: x.float().max() if False else torch.max(x)); print(f"max bf16 (read only): {nb/t/1e6:.0f} GB/s")
t=timed(lambda: y.fill_(1.0)); print(f"fill (write only): {nb/t/1e6:.0f} GB/s")
t=timed(lambda: torch.add(x, 1.0, out=y)); print(f"add scalar (read+write): {2*nb/t/1e6:.0f} GB/s")
