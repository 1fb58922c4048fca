"""The training step behind Agent.train / selfplay's train_model_job / supervised.py (agent.py:64-89,
model.py:68-72, 86-99, netencoder.py:137-181) in PyTorch, which BASELINE.json's north_star allows for the
training step.  Inputs are built on the GPU: every ply of every game is encoded by the CUDA encode kernel
(crl_encode) from the game's per-ply records.

Keras semantics kept: Adam(lr=0.002, eps=1e-7); loss = categorical cross-entropy(one-hot of the move played)
+ mean squared error(value, white-point-of-view result) + l2(0.01) on every conv / dense kernel; BatchNorm in
training mode with momentum 0.99 / eps 1e-3; a batch = `batch_size` games, one sample per ply; each game's
planes are rotated by 180 degrees with probability 0.1 while the policy target is NOT (agent.py:82-84).
"""

from __future__ import annotations

import contextlib
import os
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import boards as B
from . import runtime
from .netencoder import get_uci_labels

L2 = 0.01
BN_EPS = 1e-3
_KERNEL_IDX = None

# Arithmetic of the training step.  "fp32" (default) is the parity setting -- fp32 convolutions and matmuls without TF32,
# what the loss / gradient / Adam tests pin against the fp64 restatement of the Keras definitions.  "tf32" and "bf16" are
# opt-in throughput settings (CRL_TRAIN_PRECISION, `--precision` of the CLIs) that put the convolutions on the tensor
# cores: measured on a B200, 640 positions per step: fp32 103.5 ms, tf32 12.6 ms (8.2 x), bf16 autocast 9.9 ms (10.5 x);
# the loss agrees to 1e-6 / 1e-5 relative, individual gradient tensors of the 21-layer tower deviate from a float64 step by
# up to 13 % / 39 % of their largest entry on a random-init pack (fp32: 1 %; scripts/probe/train_precision_probe.py) --
# NOT parity-grade, hence opt-in.
# "tf32x3" keeps fp32-grade numerics ON the tensor cores: every convolution operand is split into a TF32-representable
# head and its (exact) fp32 remainder, x = x_hi + x_lo, and x * w is evaluated as x_hi*w_hi + x_hi*w_lo + x_lo*w_hi with
# fp32 accumulation -- three TF32 convolutions instead of one fp32 convolution on the CUDA cores, forward and both
# backward passes (the dropped x_lo*w_lo term is below 2^-22 of the product).  See _Conv3xTF32.  Measured on a B200, 640
# positions: 41.1 ms per step (2.5 x fp32); loss equal to fp32's to 7 digits; gradients against a float64 step: worst tensor
# 2.0e-2 of its largest entry / 7.7e-3 in L2, where fp32 itself is at 9.8e-3 / 2.7e-3 (plain tf32: 1.3e-1 / 9.9e-2) -- the
# tensor cores' fp32 accumulation, not the split, sets that floor.  fp32-grade, still opt-in.
PRECISIONS = ("fp32", "tf32x3", "tf32", "bf16")


def resolve_precision(precision=None):
    precision = precision or os.environ.get("CRL_TRAIN_PRECISION", "fp32")
    if precision not in PRECISIONS:
        raise ValueError("training precision %r: expected one of %s" % (precision, ", ".join(PRECISIONS)))
    return precision


@contextlib.contextmanager
def arithmetic(precision):
    """TF32 switches (and bf16 autocast) for the duration of a training / validation pass, restored afterwards."""
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    global _CONV_SPLIT
    saved_split = _CONV_SPLIT
    torch.backends.cudnn.allow_tf32 = precision in ("tf32", "tf32x3")     # tf32x3: convolutions only, on split operands
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    _CONV_SPLIT = precision == "tf32x3"
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
            yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        _CONV_SPLIT = saved_split


_CONV_SPLIT = False


def _tf32_split(x):
    """x = hi + lo with hi representable in TF32 (10 explicit mantissa bits, round half up in magnitude) and lo the exact
    fp32 remainder."""
    hi = ((x.contiguous().view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
    return hi, x - hi


class _Conv3xTF32(torch.autograd.Function):
    """conv2d (stride 1, 'same' padding) whose three passes each run as three TF32 tensor-core convolutions on split
    operands (see PRECISIONS): fp32-grade results at roughly a third of the fp32 CUDA-core time."""

    @staticmethod
    def forward(ctx, x, w, b, padding):
        ctx.save_for_backward(x, w)
        ctx.padding = padding
        xh, xl = _tf32_split(x)
        wh, wl = _tf32_split(w)
        y = F.conv2d(xh, wh, None, padding=padding)
        y = y + F.conv2d(xh, wl, None, padding=padding)
        y = y + F.conv2d(xl, wh, None, padding=padding)
        return y + b.view(1, -1, 1, 1)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        pad = ctx.padding
        gh, gl = _tf32_split(gy)
        xh, xl = _tf32_split(x)
        wh, wl = _tf32_split(w)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gi = torch.nn.grad.conv2d_input
            gx = gi(x.shape, wh, gh, padding=pad) + gi(x.shape, wl, gh, padding=pad) + gi(x.shape, wh, gl, padding=pad)
        if ctx.needs_input_grad[1]:
            gwf = torch.nn.grad.conv2d_weight
            gw = gwf(xh, w.shape, gh, padding=pad) + gwf(xh, w.shape, gl, padding=pad) + gwf(xl, w.shape, gh, padding=pad)
        gb = gy.sum(dim=(0, 2, 3)) if ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


def _kernel_indices():
    global _KERNEL_IDX
    if _KERNEL_IDX is None:
        from .model import pack_shapes
        _KERNEL_IDX = [i for i, s in enumerate(pack_shapes()) if len(s) >= 2]
    return _KERNEL_IDX


def _bn(x, p, o, training):
    return F.batch_norm(x, p[o + 2], p[o + 3], p[o], p[o + 1], training=training, momentum=0.01, eps=BN_EPS)


def forward_logits(p, planes_nhwc, training):
    """p: list of 140 tensors (pack order).  Returns (policy logits [B,1968], value [B])."""
    x = planes_nhwc[..., :127].permute(0, 3, 1, 2).to(p[0].dtype)    # fp32 (fp64 when a probe passes double parameters)

    def conv(x, i):
        if _CONV_SPLIT:
            return _Conv3xTF32.apply(x, p[i].permute(3, 2, 0, 1), p[i + 1], p[i].shape[0] // 2)
        return F.conv2d(x, p[i].permute(3, 2, 0, 1), p[i + 1], padding=p[i].shape[0] // 2)

    x = conv(x, 0)
    for blk in range(10):
        o = 2 + 12 * blk
        y = torch.relu(_bn(conv(x, o), p, o + 2, training))
        y = _bn(conv(y, o + 6), p, o + 8, training)
        x = torch.relu(x + y)
    ph = torch.relu(_bn(conv(x, 122), p, 124, training))
    ph = ph.permute(0, 2, 3, 1).reshape(ph.shape[0], -1)
    logits = ph @ p[128] + p[129]
    vh = torch.relu(_bn(conv(x, 130), p, 132, training))
    vh = vh.permute(0, 2, 3, 1).reshape(vh.shape[0], -1)
    vh = torch.relu(vh @ p[136] + p[137])
    value = torch.tanh(vh @ p[138] + p[139]).reshape(-1)
    return logits, value


_LABEL_IDS = None


def _label_ids():
    global _LABEL_IDS
    if _LABEL_IDS is None:
        _LABEL_IDS = {u: i for i, u in enumerate(get_uci_labels())}
    return _LABEL_IDS


def draw_flips(n_games, random_flips):
    """One `np.random.rand() < random_flips` per game in batch order, the draw DataGameSequence.__getitem__ makes
    (netencoder.py:168) -- same calls, so numpy's global stream advances as the reference's does."""
    return [bool(np.random.rand() < random_flips) for _ in range(n_games)]


def game_samples(games, flips):
    """Host arrays for every ply of every game, the samples DatasetGame.augment_game lists (dataset.py:21-43):
    boards [N,9], history bitboards [N,8,8] (most recent first, zero where the stack is exhausted), history lengths
    [N], policy target index [N], value target [N] (white-point-of-view result, 0 for an unfinished game), flip [N]."""
    labels = _label_ids()
    boards, hists, hlens, pol, val, flip_rows = [], [], [], [], [], []
    for g, flip in zip(games, flips):
        n = len(g._moves)
        if n == 0:
            continue
        recs = np.stack(g._records[:n + 1]).astype(np.uint64)
        ply = np.arange(n)
        idx = ply[:, None] - 1 - np.arange(8)[None, :]            # position i+1 plies back, i = 0..7
        h = recs[np.clip(idx, 0, None)][:, :, :8].copy()
        h[idx < 0] = 0
        result = g.get_result()
        boards.append(recs[:n])
        hists.append(h)
        hlens.append(np.minimum(ply, 8))
        pol.append(np.array([labels[m] for m in g._moves], dtype=np.int64))
        val.append(np.full(n, 0.0 if result is None else float(result), dtype=np.float32))
        flip_rows.append(np.full(n, bool(flip)))
    if not boards:
        z = np.zeros
        return z((0, 9), np.uint64), z((0, 8, 8), np.uint64), z(0, np.uint8), z(0, np.int64), z(0, np.float32), z(0, bool)
    return (np.concatenate(boards), np.concatenate(hists), np.concatenate(hlens).astype(np.uint8), np.concatenate(pol),
            np.concatenate(val), np.concatenate(flip_rows))


def encode_games(games, flips, device=None):
    """Planes [N,8,8,128] bf16 (crl_encode) for every ply of every game plus the targets (policy index, value), the
    batch DataGameSequence.__getitem__ builds (netencoder.py:159-181): a game's planes are rotated by 180 degrees when
    its flip is set, its policy target is not (reference quirk, kept)."""
    eng = runtime.scalar_engine()
    boards, hists, hlens, pol, val, flip_rows = game_samples(games, flips)
    n = boards.shape[0]
    if n == 0:
        return (torch.zeros((0, 8, 8, 128), dtype=torch.bfloat16, device=eng.device),
                torch.zeros(0, dtype=torch.int64, device=eng.device), torch.zeros(0, device=eng.device))
    bt = eng.boards_to_device(boards)
    ht = torch.from_numpy(np.ascontiguousarray(hists.transpose(1, 2, 0)).view(np.int64)).to(eng.device)
    lt = torch.from_numpy(hlens).to(eng.device)
    planes = eng.encode(bt, ht, lt)
    if flip_rows.any():
        fr = torch.from_numpy(flip_rows).to(eng.device)
        planes = torch.where(fr.view(n, 1, 1, 1), planes.flip(1, 2), planes)      # np.rot90(k=2) over (H, W)
    return planes, torch.from_numpy(pol).to(eng.device), torch.from_numpy(val).to(eng.device)


class KerasAdam:
    """tf.keras.optimizers.Adam(lr=0.002) (model.py:69) step for step: beta_1 0.9, beta_2 0.999, epsilon 1e-7 applied
    as the paper's "epsilon hat":  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  w -= lr_t * m / (sqrt(v) + eps).
    (torch.optim.Adam puts eps next to sqrt(v_hat) instead, which differs where gradients are tiny.)"""

    def __init__(self, params, lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    @torch.no_grad()
    def step(self, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        torch._foreach_mul_(self.m, self.b1)
        torch._foreach_add_(self.m, grads, alpha=1.0 - self.b1)
        torch._foreach_mul_(self.v, self.b2)
        torch._foreach_addcmul_(self.v, grads, grads, value=1.0 - self.b2)
        denom = torch._foreach_sqrt(self.v)
        torch._foreach_add_(denom, self.eps)
        torch._foreach_addcdiv_(self.params, self.m, denom, value=-lr_t)


def loss_terms(params, planes, pol, val, training=True):
    """The compiled Keras loss (model.py:68-72 + kernel_regularizer='l2' on every conv / dense layer):
    mean categorical cross-entropy(one-hot of the move played) + mean squared error(value, result) + 0.01 * sum w^2.
    Returns (total, policy_loss, value_loss, regulariser, logits)."""
    logits, value = forward_logits(params, planes, training=training)
    loss_p = F.cross_entropy(logits, pol)
    loss_v = F.mse_loss(value, val)
    reg = sum((params[i] ** 2).sum() for i in _kernel_indices()) * L2
    return loss_p + loss_v + reg, loss_p, loss_v, reg, logits


def trainable_indices():
    """Everything but the BatchNorm moving statistics (those are updated by the forward pass, momentum 0.99)."""
    from .model import pack_shapes
    return [i for i, s in enumerate(pack_shapes()) if not (len(s) == 1 and _is_bn_running(i))]


def train_step(params, opt, planes, pol, val):
    """One fit_generator batch: forward in training mode (BatchNorm batch statistics, moving statistics updated in
    place), backward, Adam.  Returns the history record of the batch."""
    total, loss_p, loss_v, reg, logits = loss_terms(params, planes, pol, val, training=True)
    grads = torch.autograd.grad(total, opt.params)
    opt.step(list(grads))
    acc = (logits.argmax(1) == pol).float().mean().item()
    return {"loss": total.item(), "policy_loss": loss_p.item(), "value_loss": loss_v.item(), "reg": reg.item(),
            "policy_acc": acc}


def validate(params, val_games, batch_size, dev):
    """The validation generator of Agent.train (agent.py:73-75): `batch_size` games per batch, no flips; Keras
    averages the per-batch means over the batches."""
    bs = max(1, min(batch_size, len(val_games)))
    sums, n = {"val_loss": 0.0, "val_policy_loss": 0.0, "val_value_loss": 0.0, "val_policy_acc": 0.0}, 0
    with torch.no_grad():
        for b in range(len(val_games) // bs):
            batch = val_games[b * bs:(b + 1) * bs]
            planes, pol, val = encode_games(batch, [False] * len(batch), dev)
            total, loss_p, loss_v, _, logits = loss_terms(params, planes, pol, val, training=False)
            sums["val_loss"] += total.item()
            sums["val_policy_loss"] += loss_p.item()
            sums["val_value_loss"] += loss_v.item()
            sums["val_policy_acc"] += (logits.argmax(1) == pol).float().mean().item()
            n += 1
    return {k: v / max(n, 1) for k, v in sums.items()}


def train(model, dataset, epochs=1, logdir=None, batch_size=1, validation_split=0, verbose=True, precision=None):
    precision = resolve_precision(precision)
    eng = runtime.scalar_engine()
    dev = eng.device
    games = list(dataset.games)
    val_games = []
    if validation_split > 0:
        split = len(games) - int(validation_split * len(games))
        games, val_games = games[:split], games[split:]
    # default: fp32 like the reference's CPU TensorFlow -- no TF32 convolutions / matmuls inside the training step
    with arithmetic(precision):
        params = [torch.tensor(w, device=dev, requires_grad=False) for w in model.weights]
        trainable = []
        for i in trainable_indices():
            params[i].requires_grad_(True)
            trainable.append(params[i])
        opt = KerasAdam(trainable, lr=0.002, epsilon=1e-7)
        bs = max(1, min(batch_size, len(games))) if games else 1
        history = []
        for ep in range(epochs):
            order = list(range(len(games) // bs))
            random.shuffle(order)                    # fit_generator shuffles a Sequence's batch order every epoch
            for b in order:
                batch = games[b * bs:(b + 1) * bs]
                flips = draw_flips(len(batch), 0.1)                      # agent.py:82-84: random_flips=.1
                planes, pol, val = encode_games(batch, flips, dev)
                rec = train_step(params, opt, planes, pol, val)
                rec.update({"epoch": ep, "batch": b})
                history.append(rec)
                if verbose:
                    print("epoch %d batch %d loss %.4f (policy %.4f value %.4f) acc %.3f" %
                          (ep, b, rec["loss"], rec["policy_loss"], rec["value_loss"], rec["policy_acc"]))
            if val_games:
                rec = validate(params, val_games, batch_size, dev)
                rec["epoch"] = ep
                history.append(rec)
                if verbose:
                    print("epoch %d val_loss %.4f (policy %.4f value %.4f) val_acc %.3f" %
                          (ep, rec["val_loss"], rec["val_policy_loss"], rec["val_value_loss"], rec["val_policy_acc"]))
    model.weights = [p.detach().cpu().numpy().astype(np.float32) for p in params]
    if logdir is not None:
        import json
        import os
        os.makedirs(logdir, exist_ok=True)
        with open(os.path.join(logdir, "train_log.jsonl"), "a") as f:
            for h in history:
                f.write(json.dumps(h) + "\n")
    return history


def _is_bn_running(i):
    """True for moving_mean / moving_var tensors of the weight pack."""
    if 2 <= i < 122:
        j = (i - 2) % 6
        return j in (4, 5)
    if 124 <= i < 128:
        return i - 124 in (2, 3)
    if 132 <= i < 136:
        return i - 132 in (2, 3)
    return False
