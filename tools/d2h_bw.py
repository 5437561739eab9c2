import torch, time
dev=torch.device('cuda:0')
for mb in (18, 72, 144, 288):
    n=mb*1024*1024//4
    d=torch.zeros(n,dtype=torch.float32,device=dev)
    h=torch.zeros(n,dtype=torch.float32).pin_memory()
    print(mb,'MB pinned?',h.is_pinned())
    s=torch.cuda.Stream(dev)
    for rep in range(3):
        torch.cuda.synchronize()
        t=time.perf_counter()
        with torch.cuda.stream(s):
            h.copy_(d,non_blocking=True)
        s.synchronize()
        dt=time.perf_counter()-t
        print('   %.1f GB/s'%(mb/1024/dt))
