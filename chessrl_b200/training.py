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

import numpy as np
import torch
import torch.nn.functional as F

from . import boards as B
from . import runtime
from .netencoder import get_uci_labels

L2 = 0.01
BN_EPS = 1e-3
_KERNEL_IDX = None


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
    x = planes_nhwc[..., :127].permute(0, 3, 1, 2).float()

    def conv(x, i):
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


def encode_games(games, flips, device):
    """Planes [N,8,8,128] bf16 for every ply of every game, plus targets."""
    eng = runtime.scalar_engine()
    labels = {u: i for i, u in enumerate(get_uci_labels())}
    boards, hists, hlens, pol, val, flip_rows = [], [], [], [], [], []
    for g, flip in zip(games, flips):
        result = g.get_result()
        recs = g._records
        for ply, mv in enumerate(g._moves):
            boards.append(recs[ply])
            prev = recs[:ply][::-1][:8]
            h = np.zeros((8, 8), dtype=np.uint64)
            for i, r in enumerate(prev):
                h[i] = r[:8]
            hists.append(h)
            hlens.append(len(prev))
            pol.append(labels[mv])
            val.append(0.0 if result is None else float(result))
            flip_rows.append(flip)
    n = len(boards)
    bt = eng.boards_to_device(np.stack(boards))
    ht = torch.from_numpy(np.ascontiguousarray(np.stack(hists).transpose(1, 2, 0)).view(np.int64)).to(eng.device)
    lt = torch.tensor(hlens, dtype=torch.uint8, device=eng.device)
    planes = eng.encode(bt, ht, lt)
    fr = torch.tensor(flip_rows, dtype=torch.bool, device=eng.device)
    if fr.any():
        planes = torch.where(fr.view(n, 1, 1, 1), planes.flip(1, 2), planes)      # np.rot90(k=2) over (H, W)
    return planes, torch.tensor(pol, device=eng.device), torch.tensor(val, dtype=torch.float32, device=eng.device)


def train(model, dataset, epochs=1, logdir=None, batch_size=1, validation_split=0, verbose=True):
    eng = runtime.scalar_engine()
    dev = eng.device
    games = list(dataset.games)
    val_games = []
    if validation_split > 0:
        split = len(games) - int(validation_split * len(games))
        games, val_games = games[:split], games[split:]
    params = [torch.tensor(w, device=dev, requires_grad=False) for w in model.weights]
    from .model import pack_shapes
    shapes = pack_shapes()
    trainable = []
    for i, s in enumerate(shapes):
        is_bn_stat = len(s) == 1 and _is_bn_running(i)
        if not is_bn_stat:
            params[i].requires_grad_(True)
            trainable.append(params[i])
    opt = torch.optim.Adam(trainable, lr=0.002, eps=1e-7)
    kidx = _kernel_indices()
    bs = max(1, min(batch_size, len(games))) if games else 1
    history = []
    for ep in range(epochs):
        for b in range(len(games) // bs):
            batch = games[b * bs:(b + 1) * bs]
            flips = [np.random.rand() < 0.1 for _ in batch]
            planes, pol, val = encode_games(batch, flips, dev)
            logits, value = forward_logits(params, planes, training=True)
            loss_p = F.cross_entropy(logits, pol)
            loss_v = F.mse_loss(value, val)
            reg = sum((params[i] ** 2).sum() for i in kidx) * L2
            loss = loss_p + loss_v + reg
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            acc = (logits.argmax(1) == pol).float().mean().item()
            history.append({"epoch": ep, "batch": b, "loss": loss.item(), "policy_loss": loss_p.item(),
                            "value_loss": loss_v.item(), "policy_acc": acc})
            if verbose:
                print("epoch %d batch %d loss %.4f (policy %.4f value %.4f) acc %.3f" %
                      (ep, b, loss.item(), loss_p.item(), loss_v.item(), acc))
        if val_games:
            with torch.no_grad():
                planes, pol, val = encode_games(val_games, [False] * len(val_games), dev)
                logits, value = forward_logits(params, planes, training=False)
                history.append({"epoch": ep, "val_policy_loss": F.cross_entropy(logits, pol).item(),
                                "val_value_loss": F.mse_loss(value, val).item()})
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
