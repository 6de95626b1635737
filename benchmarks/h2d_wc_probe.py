"""H2D rate of 1.97 GB from ordinary pinned host memory against write-combined pinned memory (cudaHostAllocWriteCombined)."""
import ctypes, torch, time
rt = ctypes.CDLL("libcudart.so.12")
n = 1024 * 250 * 1920 * 4
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def rate(ptr, label):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n), 1, ctypes.c_void_p(st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n), 1, ctypes.c_void_p(st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(label, round(ms, 3), "ms", round(n / ms / 1e6, 2), "GB/s")
for flags, label in ((0, "pinned default"), (4, "pinned write-combined"), (1, "pinned portable")):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    ctypes.memset(p, 1, n)
    rate(p.value, label)
    rt.cudaFreeHost(p)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
rate(h.data_ptr(), "torch pin_memory")
