"""Device parameter blocks for the fused kernels, exposed as ordinary ``nn.Module`` objects.

A :class:`DeviceNet` owns one fp32 parameter block ``p`` in HBM laid out for the kernels
(``W[out_pad][in_pad]`` row-major with zero padding to 16 B, bias, extra vector params), the transposed mirror
``pt`` the forward passes stream through TMA, and (if trainable) Adam ``m``/``v`` and a gradient block.
The reference-facing ``nn.Module`` (same attribute names as the reference's ``Actor``/``Critic``/``MLP``) gets
``nn.Parameter`` *views* into ``p`` so ``state_dict()`` / ``load_state_dict()`` keep the checked-in checkpoint
schema with zero copies.
"""
import torch
import torch.nn as nn

from . import _lib


def pad4(x):
    return (x + 3) & ~3


def wt_ld(out_pad):
    """row stride of the transposed mirror: the library decides (engine.cuh::wt_ld_of — bank-conflict-free for its GEMM
    microkernels), the host only lays the block out accordingly"""
    return int(_lib.lib().frl_wt_ld(int(out_pad)))


class DeviceNet:
    def __init__(self, layer_dims, device, trainable=True, x_len=0):
        """``layer_dims``: list of (in, out) in layer order (twin critics: 6 layers, head h = layers 3h..3h+2)."""
        self.device = torch.device(device)
        self.layers = []
        off, toff = 0, 0
        for (i, o) in layer_dims:
            ip, op = pad4(i), pad4(o)
            d = dict(in_=i, out=o, in_pad=ip, out_pad=op, w_off=off, b_off=off + op * ip, wt_off=toff)
            off += op * ip + op
            toff += ip * wt_ld(op) + op
            self.layers.append(d)
        self.x_off, self.x_len = off, x_len
        off += pad4(x_len)
        self.n_p, self.n_pt = off, toff
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
        self.p, self.pt = z(self.n_p), z(self.n_pt)
        self.trainable = trainable
        self.m = z(self.n_p) if trainable else None
        self.v = z(self.n_p) if trainable else None
        self.g = z(self.n_p) if trainable else None
        self._c = None

    # ---- views -------------------------------------------------------------------------------------
    def weight(self, li):
        L = self.layers[li]
        return self.p[L["w_off"]:L["w_off"] + L["out_pad"] * L["in_pad"]].view(L["out_pad"], L["in_pad"])[:L["out"], :L["in_"]]

    def bias(self, li):
        L = self.layers[li]
        return self.p[L["b_off"]:L["b_off"] + L["out"]]

    def extra(self):
        return self.p[self.x_off:self.x_off + self.x_len]

    def _state_like(self, buf, li=None):
        """view of another block (m / v / g) with the same geometry as weight(li)"""
        L = self.layers[li]
        return buf[L["w_off"]:L["w_off"] + L["out_pad"] * L["in_pad"]].view(L["out_pad"], L["in_pad"])[:L["out"], :L["in_"]]

    # ---- C descriptor ------------------------------------------------------------------------------
    def c_struct(self):
        if self._c is None:
            n = _lib.Net()
            n.p, n.pt = self.p.data_ptr(), self.pt.data_ptr()
            n.m = self.m.data_ptr() if self.trainable else None
            n.v = self.v.data_ptr() if self.trainable else None
            n.g = self.g.data_ptr() if self.trainable else None
            n.n_p, n.n_pt, n.n_layers = self.n_p, self.n_pt, len(self.layers)
            n.x_off, n.x_len = self.x_off, self.x_len
            for i, L in enumerate(self.layers):
                for k, v in L.items():
                    setattr(n.L[i], k, v)
            self._c = n
        return self._c

    def sync_mirror(self):
        """pt <- transpose(p): call after writing parameters from the host side (init / load_state_dict)."""
        import ctypes
        _lib.check(_lib.lib().frl_net_sync_mirror(ctypes.byref(self.c_struct()), _lib.stream_ptr(self.device)), "frl_net_sync_mirror")

    def copy_from(self, other):
        self.p.copy_(other.p)
        self.pt.copy_(other.pt)


class _Shim(nn.Module):
    """nn.Module whose parameters alias a DeviceNet.  Holds no compute: policy inference goes through
    ``select_action`` / ``evaluate_action`` (the batched CUDA kernel)."""

    def _bind(self, net: DeviceNet, names, extra_name=None):
        self._net = net
        for li, name in enumerate(names):
            lin = nn.Module()
            lin.weight = nn.Parameter(net.weight(li), requires_grad=False)
            lin.bias = nn.Parameter(net.bias(li), requires_grad=False)
            self.add_module(name, lin)
        if extra_name:
            self.register_parameter(extra_name, nn.Parameter(net.extra().view(1, -1), requires_grad=False))
        self.register_load_state_dict_post_hook(lambda module, incompatible: net.sync_mirror())

    def forward(self, *a, **k):
        raise RuntimeError("freerl_b200 modules are parameter containers; use select_action()/evaluate_action()")


def bind_module(net: DeviceNet, torch_module: nn.Module, names, extra_name=None):
    """Copy the freshly-initialised reference-style torch module into ``net`` (keeps torch's RNG consumption
    identical to the reference constructors) and return a shim module aliasing the device block.
    ``_Shim.__init__`` registers the extra parameter FIRST so the ``state_dict`` key order matches the reference
    (``log_std`` precedes ``l1.*``)."""
    shim = _Shim()
    if extra_name:
        shim.register_parameter(extra_name, nn.Parameter(net.extra().view(1, -1), requires_grad=False))
    shim._net = net
    with torch.no_grad():
        for li, name in enumerate(names):
            src = getattr(torch_module, name)
            net.weight(li).copy_(src.weight.detach().to(net.device))
            net.bias(li).copy_(src.bias.detach().to(net.device))
            lin = nn.Module()
            lin.weight = nn.Parameter(net.weight(li), requires_grad=False)
            lin.bias = nn.Parameter(net.bias(li), requires_grad=False)
            shim.add_module(name, lin)
        if extra_name:
            net.extra().copy_(getattr(torch_module, extra_name).detach().reshape(-1).to(net.device))
    shim.register_load_state_dict_post_hook(lambda module, incompatible: net.sync_mirror())
    net.sync_mirror()
    return shim


def alias_module(net: DeviceNet, names, extra_name=None):
    """Shim over an already-populated net (targets created by block copy, mirroring ``deepcopy``)."""
    shim = _Shim()
    if extra_name:
        shim.register_parameter(extra_name, nn.Parameter(net.extra().view(1, -1), requires_grad=False))
    shim._net = net
    for li, name in enumerate(names):
        lin = nn.Module()
        lin.weight = nn.Parameter(net.weight(li), requires_grad=False)
        lin.bias = nn.Parameter(net.bias(li), requires_grad=False)
        shim.add_module(name, lin)
    shim.register_load_state_dict_post_hook(lambda module, incompatible: net.sync_mirror())
    return shim


class _Linear3(nn.Module):
    """Reference-style construction order helper: builds nn.Linear layers exactly like the reference modules so the
    torch global RNG is consumed identically (kaiming-uniform weight, uniform bias, per layer, in order)."""

    def __init__(self, specs):
        super().__init__()
        for name, i, o in specs:
            setattr(self, name, nn.Linear(i, o))
