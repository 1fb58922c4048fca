"""Probe: the training step (chessrl_b200/training.py) at fp32 (the parity setting), tf32x3 (three TF32 tensor-core
convolutions on split operands per convolution pass), with plain TF32 convolutions / matmuls, and under bf16 autocast:
time per step and the deviation of loss / gradients from the SAME step in float64 on the same batch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chessrl_b200 import model, training  # noqa: E402

dev = torch.device("cuda")
pack = model.random_pack(0, perturb_bn=True)
N = 640
g = torch.Generator(device=dev).manual_seed(3)
planes = (torch.rand((N, 8, 8, 128), device=dev, generator=g) < 0.15).to(torch.bfloat16)
planes[..., 127] = 0
pol = torch.randint(0, 1968, (N,), device=dev, generator=g)
val = torch.randint(-1, 2, (N,), device=dev, generator=g).float()


def fresh():
    params = [torch.tensor(w, device=dev) for w in pack]
    tr = []
    for i in training.trainable_indices():
        params[i].requires_grad_(True)
        tr.append(params[i])
    return params, tr


def grads_of(mode):
    params, tr = fresh()
    with training.arithmetic(mode):
        total, lp, lv, reg, _ = training.loss_terms(params, planes, pol, val, training=True)
        gr = torch.autograd.grad(total, tr)
    return float(total.detach()), float(lp.detach()), float(lv.detach()), [x.float() for x in gr]


def grads_fp64():
    """The same step in float64 on the GPU: the yardstick (fp32 itself is only good to ~1e-2 of a gradient tensor's largest
    entry on this 21-layer tower with batch statistics)."""
    params = [torch.tensor(w, device=dev).double() for w in pack]
    tr = []
    for i in training.trainable_indices():
        params[i].requires_grad_(True)
        tr.append(params[i])
    total, lp, lv, reg, _ = training.loss_terms(params, planes, pol, val.double(), training=True)
    gr = torch.autograd.grad(total, tr)
    return float(total.detach()), float(lp.detach()), float(lv.detach()), [x.double() for x in gr]


ref = grads_fp64()
if len(sys.argv) > 1 and sys.argv[1] == "benchmark":
    torch.backends.cudnn.benchmark = True
    print("torch.backends.cudnn.benchmark = True (cuDNN picks the fastest algorithm per convolution shape by timing)")
for mode in ("fp32", "tf32x3", "tf32", "bf16"):
    t, lp, lv, gr = grads_of(mode)
    gr = [x.double() for x in gr]
    # (the bias of a convolution that feeds a BatchNorm has a mathematically zero gradient: what fp32 reports for it is
    #  rounding noise, so tensors whose reference gradient is below 1e-4 of the largest one are compared on that scale)
    gmax = max(float(b.abs().max()) for b in ref[3])
    live = [(a, b) for a, b in zip(gr, ref[3]) if float(b.abs().max()) > 1e-4 * gmax]
    dead = [(a, b) for a, b in zip(gr, ref[3]) if float(b.abs().max()) <= 1e-4 * gmax]
    worst_max = max(float((a - b).abs().max() / b.abs().max()) for a, b in live)
    worst_l2 = max(float((a - b).norm() / b.norm()) for a, b in live)
    dead_abs = max([float((a - b).abs().max()) / gmax for a, b in dead] or [0.0])
    print("   %d live tensors, %d with (near-)zero gradients: their worst |error| / largest gradient = %.2e" % (len(live), len(dead), dead_abs))
    params, tr = fresh()
    opt = training.KerasAdam(tr)

    def step():
        with training.arithmetic(mode):
            total, _, _, _, _ = training.loss_terms(params, planes, pol, val, training=True)
            opt.step(list(torch.autograd.grad(total, tr)))

    for _ in range(3):
        step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("%s: %.1f ms per %d-position step = %.0f positions/s; loss %.6f (policy %.6f value %.6f) vs fp64 %.6f; "
          "gradients vs fp64: worst tensor max-error / max %.2e, L2 %.2e" %
          (mode, ms, N, N / ms * 1e3, t, lp, lv, ref[0], worst_max, worst_l2), flush=True)
