import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from piper_b200 import engine
def ref(x,w,b,dil,slope,resid):
    xt=torch.from_numpy(x).double()
    if slope: xt=F.leaky_relu(xt,slope)
    y=F.conv1d(xt,torch.from_numpy(w).double(),None if b is None else torch.from_numpy(b).double(),dilation=dil,padding=dil*(w.shape[2]-1)//2)
    if resid is not None: y=y+torch.from_numpy(resid).double()
    return y.numpy()
cases=[(4,32,32,7,12,40000,1),(2,64,64,5,6,45001,1),(2,128,128,7,3,20003,1),(2,192,384,5,1,6001,2),(1,192,768,3,1,5000,2),(2,256,256,3,1,9999,1),(1,32,32,3,1,40000,1)]
for (B,ci,co,k,dil,L,be) in cases:
    rng=np.random.default_rng(1)
    x=rng.standard_normal((B,ci,L)).astype(np.float32)*2
    w=(rng.standard_normal((co,ci,k))/np.sqrt(ci*k)).astype(np.float32)
    b=rng.standard_normal(co).astype(np.float32)
    r=rng.standard_normal((B,co,L)).astype(np.float32)
    y=engine.debug_conv1d(be,x,w,b,dil,0.1,r)
    e=np.abs(y-ref(x,w,b,dil,0.1,r))
    bad=np.argwhere(e>3e-4)
    print((B,ci,co,k,dil,L,be),'max err',e.max(),'n bad',len(bad), 'first bad',bad[:3].tolist(), 'last bad', bad[-3:].tolist())
    if len(bad):
        ts=np.unique(bad[:,2]); print('   bad t range',ts.min(),ts.max(),'count',len(ts),'t%256 hist',np.bincount(ts%256,minlength=256).nonzero()[0][:20], 'batches',np.unique(bad[:,0]), 'channels',np.unique(bad[:,1])[:10])
