"""CPU study (PyTorch, no GPU): would Winograd convolutions keep the network inside the parity tolerance?

The residual tower is power-bound on B200, so the remaining lever is doing fewer FLOPs.  Winograd
F(2x2,3x3) needs 16 multiplies per 2x2 output tile instead of 36 (2.25x fewer tensor-core FLOPs in 20 of the 21
convolutions); F(4x4,3x3) needs 36 instead of 144 (4x fewer).  On tensor cores both operands of the transformed-domain
GEMMs are bf16: the transformed weights G g G^T and -- the new rounding step -- the transformed inputs B^T d B.
This script emulates that arithmetic exactly where it matters (operands rounded to bf16, fp32 accumulation, bf16
activations between layers, as the tcgen05 tower does today) and reports policy / value errors against the fp32
network, next to the error of today's direct bf16 convolution.  Tolerance of the GPU parity test: policy <= 2e-3,
value <= 2e-2 (tests/test_gpu_net.py)."""
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import chessrl_oracle as O  # noqa: E402
import model_torch  # noqa: E402
from chessrl_b200 import model  # noqa: E402

BN_EPS = 1e-3


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


MATS = {
    2: (torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32),          # B^T
        torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32),                  # G
        torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)),                                      # A^T
    4: (torch.tensor([[4, 0, -5, 0, 1, 0], [0, -4, -4, 1, 1, 0], [0, 4, -4, -1, 1, 0], [0, -2, -1, 2, 1, 0],
                      [0, 2, -1, -2, 1, 0], [0, 4, 0, -5, 0, 1]], dtype=torch.float32),
        torch.tensor([[1 / 4, 0, 0], [-1 / 6, -1 / 6, -1 / 6], [-1 / 6, 1 / 6, -1 / 6], [1 / 24, 1 / 12, 1 / 6],
                      [1 / 24, -1 / 12, 1 / 6], [0, 0, 1]], dtype=torch.float32),
        torch.tensor([[1, 1, 1, 1, 1, 0], [0, 1, -1, 2, -2, 0], [0, 1, 1, 4, 4, 0], [0, 1, -1, 8, -8, 1]],
                     dtype=torch.float32)),
}


def conv_winograd(x, k_hwio, bias, m):
    """x [B,C,8,8] holding bf16 values; 'same' 3x3 convolution through F(m x m, 3x3) with bf16 GEMM operands."""
    Bt, G, At = MATS[m]
    a = m + 2
    w = torch.as_tensor(k_hwio, dtype=torch.float32).permute(3, 2, 0, 1)                 # [O,I,3,3]
    U = bf(torch.einsum("ab,oibc,dc->oiad", G, w, G))                                    # transformed weights [O,I,a,a]
    xp = F.pad(x, (1, 1, 1, 1))
    tiles = xp.unfold(2, a, m).unfold(3, a, m)                                          # [B,C,T,T,a,a]
    V = bf(torch.einsum("ab,nitubc,dc->nituad", Bt, tiles, Bt))                         # transformed inputs, rounded
    M = torch.einsum("oiad,nituad->notuad", U, V)                                       # fp32 accumulate over channels
    Y = torch.einsum("ab,notubc,dc->notuad", At, M, At)                                 # [B,O,T,T,m,m]
    n, o, t = Y.shape[0], Y.shape[1], Y.shape[2]
    Y = Y.permute(0, 1, 2, 4, 3, 5).reshape(n, o, t * m, t * m)
    return Y + torch.as_tensor(bias, dtype=torch.float32).view(1, -1, 1, 1)


def conv_direct(x, k_hwio, bias):
    w = bf(torch.as_tensor(k_hwio, dtype=torch.float32)).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, w, torch.as_tensor(bias, dtype=torch.float32), padding=1)


def bn(x, g, b, mu, var):
    g, b, mu, var = (torch.as_tensor(t, dtype=torch.float32).view(1, -1, 1, 1) for t in (g, b, mu, var))
    return (x - mu) / torch.sqrt(var + BN_EPS) * g + b


def forward(pack, planes, conv):
    x = torch.as_tensor(planes, dtype=torch.float32)[..., :127].permute(0, 3, 1, 2).contiguous()
    x = bf(conv_direct(x, pack[0], pack[1]))                 # the 127-channel stem stays a direct convolution
    for blk in range(10):
        o = 2 + 12 * blk
        y = bf(torch.relu(bn(conv(x, pack[o], pack[o + 1]), *pack[o + 2:o + 6])))
        y = bn(conv(y, pack[o + 6], pack[o + 7]), *pack[o + 8:o + 12])
        x = bf(torch.relu(x + y))
    # heads in fp32 from the bf16 tower output (as model_torch does)
    dev = "cpu"
    p = model_torch._conv(x, pack[122], pack[123], dev)
    p = torch.relu(model_torch._bn(p, *pack[124:128], dev))
    p = p.permute(0, 2, 3, 1).reshape(p.shape[0], -1)
    policy = torch.softmax(p @ torch.as_tensor(pack[128]) + torch.as_tensor(pack[129]), dim=-1)
    v = model_torch._conv(x, pack[130], pack[131], dev)
    v = torch.relu(model_torch._bn(v, *pack[132:136], dev))
    v = v.permute(0, 2, 3, 1).reshape(v.shape[0], -1)
    v = torch.relu(v @ torch.as_tensor(pack[136]) + torch.as_tensor(pack[137]))
    v = torch.tanh(v @ torch.as_tensor(pack[138]) + torch.as_tensor(pack[139])).reshape(-1)
    return policy, v


def positions(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        g = O.OGame()
        for _ in range(rng.randrange(0, 60)):
            ms = g.get_legal_moves()
            if not ms or g.get_result() is not None:
                break
            g.move(rng.choice(ms))
        out.append(O.planes(g).astype(np.float32))
    return np.stack(out)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    planes = positions(int(os.environ.get("N", "48")), 3)
    # self-check of the transforms in fp64-free fp32: Winograd == direct up to fp32 rounding when nothing is rounded to bf16
    x = torch.randn(2, 8, 8, 8)
    k = np.random.default_rng(0).normal(size=(3, 3, 8, 5)).astype(np.float32)
    global bf
    keep, bf = bf, (lambda t: t)
    ref = F.conv2d(x, torch.as_tensor(k).permute(3, 2, 0, 1), None, padding=1)
    for m in (2, 4):
        err = (conv_winograd(x, k, np.zeros(5, np.float32), m) - ref).abs().max().item()
        assert err < 1e-3, (m, err)
    bf = keep
    for name, pack in (("random init (the benchmark network)", model.random_pack(0)),
                       ("perturbed BatchNorm / biases (trained-like)", model.random_pack(1, perturb_bn=True))):
        with torch.no_grad():
            p0, v0 = model_torch.forward(pack, planes)
            rows = [("direct bf16 (today's tower)", forward(pack, planes, conv_direct)),
                    ("Winograd F(2x2,3x3), bf16 operands", forward(pack, planes, lambda x, k, b: conv_winograd(x, k, b, 2))),
                    ("Winograd F(4x4,3x3), bf16 operands", forward(pack, planes, lambda x, k, b: conv_winograd(x, k, b, 4)))]
        print("== %s, %d positions; tolerance policy 2e-3 / value 2e-2" % (name, len(planes)))
        for label, (p, v) in rows:
            same = (p.argmax(1) == p0.argmax(1)).float().mean().item()
            print("  %-38s policy max|d| %.2e  value max|d| %.2e  mean|d| %.2e  argmax agreement %.3f" % (
                label, (p - p0).abs().max().item(), (v - v0).abs().max().item(), (v - v0).abs().mean().item(), same))


if __name__ == "__main__":
    main()
