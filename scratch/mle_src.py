import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import airs as A, stark_backend_b200 as sb
from stark_backend_b200 import lib as L
import test_jit as tj
lib=L.load_library()
def mle_source(air,h,w,D,n_airs=1):
    ctx=sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, air.constraint_degree, air.need_rot, tj._Shape(h,w), air.public_values)
    keep=[]; c=ctx.c(keep)
    n=lib.swirl_jit_mle_source(C.byref(c), D, n_airs, None, 0)
    buf=C.create_string_buffer(n+1)
    lib.swirl_jit_mle_source(C.byref(c), D, n_airs, buf, n+1)
    return buf.value.decode()
if __name__=="__main__":
    air=A.benchmark(3,256,256,32,np.random.default_rng(0))
    src=mle_source(air,1<<20,256,3)
    i=src.index('SW_MLE_SIGNATURE(swirl_mle_jit)')
    j=src.index('// SUB 0')
    print(src[i:j])
    print([l for l in src[j:].split('\n') if l.startswith('// SUB')])
