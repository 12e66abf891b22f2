import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from screencounter_b200 import rcpp
from screencounter_b200._lib import lib
ctx = rcpp.context(0)
L = lib()
for mb in (1, 16, 64, 256, 1024):
    for rep in range(3):
        p = C.c_void_p()
        t0 = time.perf_counter(); L.scg_device_alloc(ctx, C.c_size_t(mb << 20), C.byref(p)); t1 = time.perf_counter()
        L.scg_device_free(ctx, p); t2 = time.perf_counter()
        print("%5d MB: alloc+memset %.2f ms, free %.2f ms" % (mb, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
