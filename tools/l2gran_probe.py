import ctypes, os, sys, torch
sys.path.insert(0, '.')
torch.zeros(1, device='cuda')
rt = ctypes.CDLL('libcudart.so.12')
v = ctypes.c_size_t()
rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print('L2 fetch granularity before', v.value)
if len(sys.argv) > 1:
    print('set ->', rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(sys.argv[1]))))
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print('after', v.value)
sys.argv = ['membench']
import runpy
try:
    runpy.run_path('tools/membench.py', run_name='__main__')
except SystemExit:
    pass
