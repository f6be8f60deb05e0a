import torch, time
dev='cuda'
h_in=torch.empty(5234688,dtype=torch.uint8).pin_memory(); d_in=torch.empty_like(h_in,device=dev)
d_out=torch.empty(15400960,dtype=torch.uint8,device=dev); h_out=torch.empty(15400960,dtype=torch.uint8).pin_memory()
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def run(k,both=True,h2d=True,d2h=True):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(k):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/k*1e6
run(5)
print('h2d only us',run(50,d2h=False),' GB/s',5.23e6/run(50,d2h=False)/1e3)
print('d2h only us',run(50,h2d=False),' GB/s',15.4e6/run(50,h2d=False)/1e3)
print('both us',run(50))
