import torch, time
N = 20161
A = torch.zeros((N, N), dtype=torch.float64, device='cuda')
t=time.perf_counter(); host = torch.empty((N, N), dtype=torch.float64).pin_memory(); print('pin alloc', time.perf_counter()-t)
for k in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); host.copy_(A); torch.cuda.synchronize(); dt = time.perf_counter()-t
    print('pinned D2H %.1f ms  %.1f GB/s' % (dt*1e3, N*N*8/dt/1e9))
pg = torch.empty((N, N), dtype=torch.float64)
torch.cuda.synchronize(); t = time.perf_counter(); pg.copy_(A); torch.cuda.synchronize(); dt = time.perf_counter()-t
print('pageable D2H %.1f ms  %.1f GB/s' % (dt*1e3, N*N*8/dt/1e9))
t=time.perf_counter(); B = torch.empty((N, N), dtype=torch.float64, device='cuda'); torch.cuda.synchronize(); print('cuda alloc', time.perf_counter()-t)
import ctypes
rt = ctypes.CDLL('libcudart.so.12')
p = ctypes.c_void_p()
t=time.perf_counter(); rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(N*N*8)); print('cudaMalloc', time.perf_counter()-t)
t=time.perf_counter(); rt.cudaMemcpy(ctypes.c_void_p(host.data_ptr()), p, ctypes.c_size_t(N*N*8), 2); print('cudaMemcpy to pinned', time.perf_counter()-t)
t=time.perf_counter(); rt.cudaFree(p); print('cudaFree', time.perf_counter()-t)
