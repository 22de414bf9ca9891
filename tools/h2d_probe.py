"""Dev tool: host->device bandwidth variants (one copy, two concurrent copies, write-combined pinned memory)."""
import time, ctypes, torch
n = 1270080000
cudart = ctypes.CDLL("libcudart.so.12") if False else torch.cuda.cudart()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
src = torch.empty(n, dtype=torch.uint8).pin_memory()
src.fill_(3)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def one(): dst.copy_(src, non_blocking=True)
print("one copy            %.2f GB/s" % (n / timed(one) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for parts in (2, 4, 8):
    def split():
        k = n // parts
        for i in range(parts):
            with torch.cuda.stream(s1 if i % 2 == 0 else s2):
                dst[i * k:(i + 1) * k].copy_(src[i * k:(i + 1) * k], non_blocking=True)
    print("%d parts, 2 streams  %.2f GB/s" % (parts, n / timed(split) / 1e9))
# write-combined pinned memory through the runtime API
lib = ctypes.CDLL("libcudart.so")
p = ctypes.c_void_p()
rc = lib.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
print("cudaHostAlloc WC rc", rc)
if rc == 0:
    ctypes.memset(p, 3, n)
    def wc(): lib.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), p, ctypes.c_size_t(n), 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    print("write-combined      %.2f GB/s" % (n / timed(wc) / 1e9))
