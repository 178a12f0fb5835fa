"""How fast can host memory be pinned?  cudaHostAlloc vs (parallel first touch + cudaHostRegister), one thread vs several."""
import ctypes as C, mmap, threading, time
import numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = torch.cuda.cudart()
MB = 1 << 20
def t_hostalloc(n):
    t0 = time.time(); x = torch.empty(n, dtype=torch.uint8, pin_memory=True); dt = time.time() - t0; return dt, x
def touch(buf, threads):
    a = np.frombuffer(buf, dtype=np.uint8); n = len(a); per = (n + threads - 1) // threads
    ts = [threading.Thread(target=lambda lo=lo: a[lo:lo + per:4096].__setitem__(slice(None), 1)) for lo in range(0, n, per)]
    [t.start() for t in ts]; [t.join() for t in ts]
def t_register(n, threads, populate=False):
    t0 = time.time()
    m = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | (mmap.MAP_POPULATE if populate else 0))
    t1 = time.time()
    if not populate: touch(m, threads)
    t2 = time.time()
    addr = C.addressof(C.c_char.from_buffer(m))
    r = rt.cudaHostRegister(addr, n, 0)
    t3 = time.time()
    rt.cudaHostUnregister(addr)
    return t1 - t0, t2 - t1, t3 - t2, int(r)
n = 512 * MB
for rep in range(2):
    dt, x = t_hostalloc(n); print(f"cudaHostAlloc 512MB: {dt*1e3:.1f} ms  {n/dt/1e9:.2f} GB/s"); del x
for th in (1, 4, 16):
    a, b, c, r = t_register(n, th); print(f"mmap {a*1e3:.1f} ms + touch x{th} {b*1e3:.1f} ms + cudaHostRegister {c*1e3:.1f} ms rc={r}  total {n/(a+b+c)/1e9:.2f} GB/s")
a, b, c, r = t_register(n, 1, True); print(f"mmap MAP_POPULATE {a*1e3:.1f} ms + cudaHostRegister {c*1e3:.1f} ms rc={r}")
# untouched memory registered directly
t0 = time.time(); m = mmap.mmap(-1, n); addr = C.addressof(C.c_char.from_buffer(m)); r = rt.cudaHostRegister(addr, n, 0); print(f"register untouched: {(time.time()-t0)*1e3:.1f} ms rc={int(r)}"); rt.cudaHostUnregister(addr)
# several threads allocating at once
def par(k, each):
    out = [None] * k
    def w(i): out[i] = t_hostalloc(each)
    t0 = time.time(); ts = [threading.Thread(target=w, args=(i,)) for i in range(k)]; [t.start() for t in ts]; [t.join() for t in ts]
    return time.time() - t0
for k in (1, 2, 4, 8):
    dt = par(k, 256 * MB); print(f"{k} threads x cudaHostAlloc 256MB: {dt*1e3:.1f} ms  {k*256*MB/dt/1e9:.2f} GB/s")
def par_reg(k, each):
    ms = [mmap.mmap(-1, each) for _ in range(k)]
    t0 = time.time()
    def w(i):
        touch(ms[i], 2); rt.cudaHostRegister(C.addressof(C.c_char.from_buffer(ms[i])), each, 0)
    ts = [threading.Thread(target=w, args=(i,)) for i in range(k)]; [t.start() for t in ts]; [t.join() for t in ts]
    dt = time.time() - t0
    for m in ms: rt.cudaHostUnregister(C.addressof(C.c_char.from_buffer(m)))
    return dt
for k in (1, 4, 8):
    dt = par_reg(k, 256 * MB); print(f"{k} threads x (touch + cudaHostRegister 256MB): {dt*1e3:.1f} ms  {k*256*MB/dt/1e9:.2f} GB/s")
