import sys, ctypes, torch
sys.path.insert(0, '.')
from freerl_b200 import _lib
from freerl_b200.nets import DeviceNet
dev = torch.device('cuda')
net = DeviceNet([(24, 128), (128, 128), (128, 4), (128, 128), (20, 128), (128, 8)], dev, trainable=False)
net.p.normal_(0, 0.05); net.sync_mirror()
sink = torch.zeros(256, device=dev)
lib = _lib.lib()
lib.frl_debug_opbench.argtypes = [ctypes.POINTER(_lib.Net), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
def run(li, mode, iters=4000):
    for _ in range(2):
        lib.frl_debug_opbench(ctypes.byref(net.c_struct()), li, 3, iters, mode, 64, ctypes.c_void_p(sink.data_ptr()), _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.frl_debug_opbench(ctypes.byref(net.c_struct()), li, 3, iters, mode, 64, ctypes.c_void_p(sink.data_ptr()), _lib.stream_ptr(dev))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
print('128->128 %.3f | 24->128 %.3f | 128->4 %.3f  (gemm only, us)' % (run(1, 0), run(0, 0), run(2, 0)))
