import torch, time
x = torch.empty(1966080000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for chunk in (1966080000, 65280000, 16<<20):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        off = 0
        while off < x.numel():
            n = min(chunk, x.numel() - off)
            d[off:off+n].copy_(x[off:off+n], non_blocking=True); off += n
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"chunk {chunk/1e6:8.1f} MB: {x.numel()/dt/1e9:.2f} GB/s  ({dt*1e3:.2f} ms)")
