import torch, time
torch.cuda.init(); torch.zeros(1).cuda()
for mb in (262, 1310):
    t=time.perf_counter(); h=torch.empty(mb*1024*1024//4, dtype=torch.float32, pin_memory=True); t1=time.perf_counter()-t
    d=torch.empty_like(h, device='cuda'); torch.cuda.synchronize()
    t=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()-t
    t=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t3=time.perf_counter()-t
    hp=torch.empty(mb*1024*1024//4, dtype=torch.float32)
    t=time.perf_counter(); hp.copy_(d); torch.cuda.synchronize(); t4=time.perf_counter()-t
    t=time.perf_counter(); a=h.numpy().copy(); t5=time.perf_counter()-t
    print(f"{mb} MB: pinned alloc {t1*1e3:.1f} ms, D2H pinned {t2*1e3:.1f}/{t3*1e3:.1f} ms ({mb/t3/1e3:.1f} GB/s), D2H pageable {t4*1e3:.1f} ms, host memcpy {t5*1e3:.1f} ms")
    del h, d, hp
