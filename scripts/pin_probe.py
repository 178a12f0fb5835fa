import time, sys, torch, numpy as np
path=sys.argv[1]
def rd(buf):
    mv=memoryview(buf)
    t0=time.time(); n=0
    with open(path,'rb',buffering=0) as f:
        while True:
            k=f.readinto(mv[:64<<20])
            if not k: break
            n+=k
    return n/(time.time()-t0)/1e9
pinned=torch.empty(64<<20,dtype=torch.uint8,pin_memory=True).numpy()
plain=np.empty(64<<20,dtype=np.uint8); plain[:]=0
print('read into plain  GB/s', rd(plain), rd(plain))
print('read into pinned GB/s', rd(pinned), rd(pinned))
from cutseq_b200 import native
L=native.lib()
for name,b in (('plain',plain),('pinned',pinned)):
    t0=time.time(); c=L.csq_count_newlines(b.ctypes.data,b.size); print('count',name,b.size/(time.time()-t0)/1e9)
