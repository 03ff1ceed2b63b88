import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piper_b200 import _lib
lib=_lib.load()
out=(C.c_uint64*2)()
print("N n_acc mode | issue cyc/iter | total cyc/iter   (mode 512: no MMA in the loop; 1024: 4 MMAs per iteration)")
for N in (32,128,256):
    for mode in (0, 512, 1024):
        for n_acc in (1,2):
            if n_acc*N>512: continue
            it=2000
            _lib.check(lib.pb200_debug_mma_bench(N,0,n_acc,it,mode,out))
            print(f"{N:4d} {n_acc} {mode:4d} | {out[0]/it:8.1f} | {out[1]/it:8.1f}")
