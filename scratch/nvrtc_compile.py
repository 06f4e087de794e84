import ctypes as C, sys, time
def compile_src(src, extra=()):
    nv = C.CDLL("libnvrtc.so.12")
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src.encode(), b"jit.cu", 0, None, None) == 0
    o = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo", b"--extra-device-vectorization", b"--ptxas-options=-v"] + list(extra)
    opts = (C.c_char_p * len(o))(*o)
    t=time.time()
    rc = nv.nvrtcCompileProgram(prog, len(o), opts)
    dt=time.time()-t
    n = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(n.value + 1)
    nv.nvrtcGetProgramLog(prog, log)
    cub=None
    if rc==0:
        nv.nvrtcGetCUBINSize(prog, C.byref(n))
        cub=C.create_string_buffer(n.value)
        nv.nvrtcGetCUBIN(prog, cub)
    return rc, log.value.decode(), dt, (cub.raw if cub else None)
if __name__=="__main__":
    sys.path.insert(0,'/root/repo/scratch')
    import mle_src as M, numpy as np
    import airs as A
    air=A.benchmark(3,256,256,32,np.random.default_rng(0))
    for D in (3,):
        src=M.mle_source(air,1<<20,256,D)
        rc,log,dt,cub=compile_src(src)
        print(D, rc, "%.1fs"%dt, len(cub or b''))
        print(log[-1500:])
        open('/tmp/mle_jit.cubin','wb').write(cub)
