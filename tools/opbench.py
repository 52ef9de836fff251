"""Scratch: time one layer op (gemm_rk / layer_fwd) in isolation on the GPU."""
import sys, ctypes
import torch
sys.path.insert(0, '.')
from freerl_b200 import _lib
from freerl_b200.nets import DeviceNet
dev = torch.device('cuda')
net = DeviceNet([(24, 128), (128, 128), (128, 4), (128, 128), (20, 128), (128, 8)], dev, trainable=False)
net.p.normal_(0, 0.05); net.sync_mirror()
sink = torch.zeros(256, device=dev)
lib = _lib.lib()
lib.frl_debug_opbench.argtypes = [ctypes.POINTER(_lib.Net), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
def run(li, li2, mode, ncta, iters=2000):
    for _ in range(2):
        lib.frl_debug_opbench(ctypes.byref(net.c_struct()), li, li2, iters, mode, ncta, ctypes.c_void_p(sink.data_ptr()), _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.frl_debug_opbench(ctypes.byref(net.c_struct()), li, li2, iters, mode, ncta, ctypes.c_void_p(sink.data_ptr()), _lib.stream_ptr(dev))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
names = {0: '24->128', 1: '128->128', 2: '128->4', 4: '20->128', 5: '128->8'}
for ncta in (1, 64, 148):
    for li in (1, 0, 2):
        print('ncta=%3d %-9s gemm only %.3f us | tma no-prefetch %.3f us | tma prefetch(alt w/ 3) %.3f us' % (
            ncta, names[li], run(li, li, 0, ncta), run(li, li, 1, ncta), run(li, 3, 2, ncta)))
