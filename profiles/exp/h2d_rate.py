import torch
dev=torch.device('cuda',0)
for nbytes in (5_950_000_000, 744_000_000, 64_000_000, 4_000_000):
    h=torch.empty(nbytes,dtype=torch.uint8).pin_memory(); d=torch.empty(nbytes,dtype=torch.uint8,device=dev)
    for _ in range(2): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    reps=max(1,int(6e9//nbytes))
    e0.record()
    for _ in range(reps): d.copy_(h,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(nbytes, reps, nbytes*reps/e0.elapsed_time(e1)/1e6,'GB/s')
    del h,d
