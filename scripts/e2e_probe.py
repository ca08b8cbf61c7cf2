import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
import riptrm_b200 as rb
from riptrm_b200 import _lib
B=4096
Zh=torch.empty((B,50,50),dtype=torch.float64).pin_memory(); x0h=torch.empty((B,50),dtype=torch.float64).pin_memory(); y0h=torch.empty((B,50),dtype=torch.float64).pin_memory()
rb.datagen.nonnegpca_batch(0,B,50,out=(Zh.numpy(),x0h.numpy(),y0h.numpy()))
xh=torch.empty((B,50),dtype=torch.float64).pin_memory(); yh=torch.empty((B,50),dtype=torch.float64).pin_memory(); smh=torch.empty((B,16),dtype=torch.float64).pin_memory()
opt=rb.options.default_option(); opt.update(TRS_solver="tCG",second_order_stationarity=False,maxiter=30,inner_maxiter=1000,tolresid=0,maxtime=1e9)
s=rb.BatchSolver.nonnegpca_from_arrays(Zh.numpy(),x0h.numpy(),y0h.numpy()); s.set_options(opt,0,0)
for rep in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    s.set_nonnegpca(Zh.numpy(), _lib.HOST)
    torch.cuda.synchronize(); t1=time.perf_counter()
    _lib.check(s.lib.riptrm_solve(s.handle.h,_lib.ptr(x0h.numpy()),_lib.ptr(y0h.numpy()),_lib.ptr(xh.numpy()),_lib.ptr(yh.numpy()),_lib.ptr(smh.numpy()),None,_lib.HOST,None))
    t2=time.perf_counter()
    print(f"set_nonnegpca {1e3*(t1-t0):.2f} ms, solve(host) {1e3*(t2-t1):.2f} ms, kernel {s.kernel_ms:.2f} ms")
