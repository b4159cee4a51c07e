import torch
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    best=None
    for _ in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        t=e0.elapsed_time(e1)/reps*1e3
        best=t if best is None else min(best,t)
    return best
for mb in (101, 400, 1600):
    n = mb*1000*1000//4
    bufs=[torch.empty(n, device="cuda") for _ in range(4)]
    i=[0]
    def f():
        bufs[i[0]%4].zero_(); i[0]+=1
    t=timed(f)
    print(f"memset {mb} MB: {t:.1f} us = {mb*1e6/t/1e3:.0f} GB/s")
    src=torch.empty(n, device="cuda")
    def c():
        bufs[i[0]%4].copy_(src); i[0]+=1
    t=timed(c)
    print(f"copy   {mb} MB -> {mb} MB: {t:.1f} us = {2*mb*1e6/t/1e3:.0f} GB/s (read+write)")
