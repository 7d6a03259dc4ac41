"""RPN classification + regression heads on the tensor-core chain kernel.

The reference runs two Conv1d stacks over the backbone features (lib/net/rpn.py:19-47, :76-77):
    rpn_cls_layer = Conv1d(C,128,bn) - Dropout - Conv1d(128,1, activation=None)
    rpn_reg_layer = Conv1d(C,128,bn) - Dropout - Conv1d(128,reg_channel, activation=None)
and transposes both results to point-major.  In eval mode (Dropout = identity, BN = running statistics) the two stacks
are ONE two-layer MLP over point-major rows: layer 0 stacks the two hidden layers (C -> 256, BN folded into the
weights, ReLU), layer 1 is block diagonal (256 -> 1 + reg_channel, linear: `prb_mlp_desc.flags & 1`).  The rows are
the point-major twin the last FP level already wrote, the output rows are already (B,N,·): no transposes, one launch.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi as C
from ..pointnet2 import pointnet2_modules as pm


def _layers(seq):
    """[(conv, bn or None, has_relu)] of a reference head stack; Dropout is the identity in eval mode"""
    out = []
    for layer in seq.children():
        if isinstance(layer, nn.Dropout):
            continue
        names = [k for k, _ in layer.named_children()]
        out.append((layer.conv, layer.bn.bn if "bn" in names else None, "activation" in names))
    return out


def _fold(conv, bn):
    W = conv.weight.detach().reshape(conv.out_channels, -1).float().cpu()
    b = conv.bias.detach().float().cpu() if conv.bias is not None else torch.zeros(conv.out_channels)
    if bn is None:
        return W, b
    inv = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)).cpu()
    return W * inv[:, None], (bn.bias.detach().float().cpu() - bn.running_mean.detach().float().cpu() * inv) + inv * b


class FusedRPNHeads:
    """cache of the packed two-layer image of (rpn_cls_layer, rpn_reg_layer); rebuilt when a parameter changes"""

    def __init__(self):
        self.key = None
        self.state = None

    @staticmethod
    def supported(cls_seq, reg_seq):
        try:
            a, b = _layers(cls_seq), _layers(reg_seq)
        except AttributeError:
            return False
        ok = len(a) == 2 and len(b) == 2 and a[0][2] and b[0][2] and not a[1][2] and not b[1][2]
        ok = ok and a[0][0].in_channels == b[0][0].in_channels and a[0][0].out_channels + b[0][0].out_channels <= 512
        ok = ok and all(bn is None or not bn.training for _, bn, _ in a + b)
        return ok and all(tuple(c.kernel_size) in ((1,), (1, 1)) for c, _, _ in a + b)

    def _build(self, cls_seq, reg_seq, device):
        (c0, cb0, _), (c1, _, _) = _layers(cls_seq)
        (r0, rb0, _), (r1, _, _) = _layers(reg_seq)
        Wc0, sc0 = _fold(c0, cb0)
        Wr0, sr0 = _fold(r0, rb0)
        Wc1, sc1 = _fold(c1, None)
        Wr1, sr1 = _fold(r1, None)
        hc, hr = Wc0.shape[0], Wr0.shape[0]
        W0 = torch.cat([Wc0, Wr0], 0).contiguous()
        W1 = torch.zeros(Wc1.shape[0] + Wr1.shape[0], hc + hr)
        W1[:Wc1.shape[0], :hc] = Wc1
        W1[Wc1.shape[0]:, hc:] = Wr1
        c_in, c_out = W0.shape[1], [W0.shape[0], W1.shape[0]]
        lib = C.lib()
        co = (ctypes.c_int * 3)(c_out[0], c_out[1], 0)
        nbytes = lib.prb_mlp_packed_bytes_ex(2, 0, 2, c_in, co)
        host = np.zeros(nbytes // 4, dtype=np.float32)
        ws = [W0.contiguous(), W1.contiguous()]
        wp = (ctypes.c_void_p * 2)(*[w.data_ptr() for w in ws])
        C.check(lib.prb_mlp_pack_weights_ex(2, 0, 2, c_in, co, wp, host.ctypes.data_as(ctypes.c_void_p)), "mlp_pack")
        pad = lambda v: torch.nn.functional.pad(v, (0, (-v.numel()) % 32))
        packed = torch.from_numpy(host).to(device)
        shift = torch.cat([pad(torch.cat([sc0, sr0])), pad(torch.cat([sc1, sr1]))]).contiguous().to(device)
        desc = C.MlpDesc()
        desc.num_layers, desc.c_in = 2, c_in
        desc.c_out[0], desc.c_out[1], desc.c_out[2] = c_out[0], c_out[1], 0
        desc.packed_w, desc.scale, desc.shift, desc.flags = packed.data_ptr(), None, shift.data_ptr(), 1
        return dict(desc=desc, keep=(packed, shift), co=co, c_in=c_in, n_cls=Wc1.shape[0], n_out=c_out[1])

    def __call__(self, cls_seq, reg_seq, features):
        """features (B,C,N) -> rpn_cls (B,N,n_cls), rpn_reg (B,N,reg_channel), as lib/net/rpn.py:76-77"""
        tensors = [p for seq in (cls_seq, reg_seq) for p in list(seq.parameters()) + list(seq.buffers())]
        key = (str(features.device),) + tuple((id(t), t._version) for t in tensors)
        if key != self.key:
            self.state, self.key = self._build(cls_seq, reg_seq, features.device), key
        st = self.state
        B, Cc, N = features.shape
        assert Cc == st["c_in"]
        rows = pm._point_major(features)            # (B,N,C): the twin written by the last FP launch, or one transpose
        lib = C.lib()
        pitch = (st["n_out"] + 31) // 32 * 32
        out = torch.empty((B * N, pitch), dtype=torch.float32, device=features.device)
        wsb = lib.prb_rows_workspace_bytes(C.c_long(B * N), Cc, 2, st["co"])
        ws = torch.empty(wsb, dtype=torch.uint8, device=features.device)
        with torch.cuda.device(features.device):
            C.check(lib.prb_mlp_rows(C.c_long(B * N), Cc, C.ptr(rows), ctypes.byref(st["desc"]), C.ptr(out), pitch, C.ptr(ws),
                                     C.c_size_t(wsb), C.stream()), "mlp_rows(heads)")
        out = out.view(B, N, pitch)
        return out[:, :, :st["n_cls"]].contiguous(), out[:, :, st["n_cls"]:st["n_out"]].contiguous()


def rpn_heads(rpn_module, backbone_features):
    """eval-mode replacement of `rpn_cls_layer(f).transpose(1,2).contiguous()`, `rpn_reg_layer(f).transpose(...)`
    for a reference lib.net.rpn.RPN instance (or anything with the same two attributes)"""
    cache = rpn_module.__dict__.setdefault("_prb_heads", FusedRPNHeads())
    fused_ok = (not torch.is_grad_enabled()) and not rpn_module.training and pm._fused_enabled() and \
        FusedRPNHeads.supported(rpn_module.rpn_cls_layer, rpn_module.rpn_reg_layer)
    if not fused_ok:
        return (rpn_module.rpn_cls_layer(backbone_features).transpose(1, 2).contiguous(),
                rpn_module.rpn_reg_layer(backbone_features).transpose(1, 2).contiguous())
    return cache(rpn_module.rpn_cls_layer, rpn_module.rpn_reg_layer, backbone_features)
