"""ORACLE / TEST INFRASTRUCTURE ONLY.  Plain PyTorch fp32 restatement of model.ChessModel's forward graph
(model.py:31-63, 111-122) evaluated from the 140-tensor weight pack (chessrl_b200/model.py documents the order).
Used as the floating-point reference for the tcgen05 network kernels and as the CPU baseline's evaluator."""

import torch
import torch.nn.functional as F

BN_EPS = 1e-3   # Keras BatchNormalization default


def _t(a, device):
    return torch.as_tensor(a, dtype=torch.float32, device=device)


def _conv(x, k, b, device):
    w = _t(k, device).permute(3, 2, 0, 1).contiguous()       # HWIO -> OIHW
    return F.conv2d(x, w, _t(b, device), padding=k.shape[0] // 2)


def _bn(x, g, b, m, v, device):
    g, b, m, v = (_t(a, device).view(1, -1, 1, 1) for a in (g, b, m, v))
    return (x - m) / torch.sqrt(v + BN_EPS) * g + b


def forward(pack, planes_nhwc, device="cpu", input_dtype=None, emulate_bf16_activations=False, taps=None):
    """planes_nhwc: [B,8,8,127] (or 128 with a zero pad channel).  Returns (policy [B,1968], value [B]).

    emulate_bf16_activations=True mirrors the rounding points of the tcgen05 path (csrc/net.cu) in an otherwise fp32
    graph: convolution kernels and the policy dense kernel rounded to bf16, activations rounded to bf16 after every
    convolution's epilogue (BatchNorm, skip, ReLU) and before the dense policy layer; accumulation, BatchNorm folding,
    the 1x1 head convolutions' weights, the value head and the softmax stay fp32.
    taps: optional dict that receives 'act' (list of the 21 convolution outputs, NHWC), 'pf' [B,128], 'vf' [B,64],
    'logits' [B,1968]."""
    x = torch.as_tensor(planes_nhwc, dtype=torch.float32, device=device)[..., :127].permute(0, 3, 1, 2).contiguous()
    emu = emulate_bf16_activations

    def q(t):
        return t.to(torch.bfloat16).to(torch.float32) if emu else t

    def qw(k):
        import numpy as np
        if not emu:
            return k
        return torch.as_tensor(np.asarray(k), dtype=torch.float32).to(torch.bfloat16).to(torch.float32).numpy()

    acts = []

    def tap(t):
        if taps is not None:
            acts.append(t.permute(0, 2, 3, 1).contiguous())
        return t

    x = tap(q(_conv(x, qw(pack[0]), pack[1], device)))
    for blk in range(10):
        o = 2 + 12 * blk
        y = _conv(x, qw(pack[o]), pack[o + 1], device)
        y = tap(q(torch.relu(_bn(y, *pack[o + 2:o + 6], device))))
        y = _conv(y, qw(pack[o + 6]), pack[o + 7], device)
        y = _bn(y, *pack[o + 8:o + 12], device)
        x = tap(q(torch.relu(x + y)))
    p = _conv(x, pack[122], pack[123], device)
    p = torch.relu(_bn(p, *pack[124:128], device))
    p = q(p.permute(0, 2, 3, 1).reshape(p.shape[0], -1))        # Keras Flatten of NHWC: (h*8+w)*2+c
    logits = p @ _t(qw(pack[128]), device) + _t(pack[129], device)
    policy = torch.softmax(logits, dim=-1)
    v = _conv(x, pack[130], pack[131], device)
    v = torch.relu(_bn(v, *pack[132:136], device))
    vf = v.permute(0, 2, 3, 1).reshape(v.shape[0], -1)
    v = torch.relu(vf @ _t(pack[136], device) + _t(pack[137], device))
    v = torch.tanh(v @ _t(pack[138], device) + _t(pack[139], device)).reshape(-1)
    if taps is not None:
        taps.update({"act": acts, "pf": p, "vf": vf, "logits": logits})
    return policy, v
