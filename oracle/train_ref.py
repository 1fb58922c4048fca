"""ORACLE / TEST INFRASTRUCTURE ONLY.  Independent float64 restatement of the reference's training step:

  model.py:68-72    compile(Adam(lr=0.002), loss=['categorical_crossentropy', 'mean_squared_error'])
  model.py:33-60    kernel_regularizer='l2' on every Conv2D / Dense  ->  + 0.01 * sum(kernel ** 2) each
  model.py:111-122  residual block with BatchNormalization(axis=-1) (Keras defaults: momentum 0.99, epsilon 1e-3)
  agent.py:64-89    one batch = the plies of `batch_size` games

Written against the Keras definitions, not against chessrl_b200/training.py: BatchNorm is spelled out (batch mean,
biased batch variance for the normalisation, moving statistics updated with momentum 0.99 -- the variance that goes
into the moving average is the UNBIASED one, as TensorFlow's fused batch-norm kernel returns it), cross-entropy is
Keras's formula on the softmax OUTPUT (renormalise, clip to [1e-7, 1 - 1e-7], -sum(target * log)), Adam is
tf.keras's update (epsilon-hat form).  TensorFlow itself cannot be installed here: parity with TF's
floating-point results is unpinned; what is pinned is the arithmetic definition.
"""

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
L2 = 0.01
KERAS_EPS = 1e-7


def _conv(x, k, b):
    return F.conv2d(x, k.permute(3, 2, 0, 1), b, padding=k.shape[0] // 2)


def _bn_train(x, gamma, beta, mean, var, new_stats):
    """x: [B,C,H,W].  Returns the normalised tensor; appends the updated (moving_mean, moving_var) to new_stats."""
    n = x.shape[0] * x.shape[2] * x.shape[3]
    mu = x.mean(dim=(0, 2, 3))
    dev = x - mu.view(1, -1, 1, 1)
    var_b = (dev * dev).mean(dim=(0, 2, 3))                      # biased: used to normalise
    y = dev / torch.sqrt(var_b + BN_EPS).view(1, -1, 1, 1) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    var_u = var_b * (n / max(n - 1, 1))                          # unbiased: goes into the moving average
    new_stats.append((BN_MOMENTUM * mean + (1 - BN_MOMENTUM) * mu.detach(),
                      BN_MOMENTUM * var + (1 - BN_MOMENTUM) * var_u.detach()))
    return y


def kernel_indices(shapes):
    return [i for i, s in enumerate(shapes) if len(s) >= 2]


def forward_train(p, planes_nhwc):
    """p: 140 float64 tensors in pack order.  Returns (softmax policy, value, updated BN statistics in graph order)."""
    x = planes_nhwc[..., :127].permute(0, 3, 1, 2)
    stats = []
    x = _conv(x, p[0], p[1])
    for blk in range(10):
        o = 2 + 12 * blk
        y = torch.relu(_bn_train(_conv(x, p[o], p[o + 1]), p[o + 2], p[o + 3], p[o + 4], p[o + 5], stats))
        y = _bn_train(_conv(y, p[o + 6], p[o + 7]), p[o + 8], p[o + 9], p[o + 10], p[o + 11], stats)
        x = torch.relu(x + y)
    ph = torch.relu(_bn_train(_conv(x, p[122], p[123]), p[124], p[125], p[126], p[127], stats))
    ph = ph.permute(0, 2, 3, 1).reshape(ph.shape[0], -1)
    policy = torch.softmax(ph @ p[128] + p[129], dim=-1)
    vh = torch.relu(_bn_train(_conv(x, p[130], p[131]), p[132], p[133], p[134], p[135], stats))
    vh = vh.permute(0, 2, 3, 1).reshape(vh.shape[0], -1)
    vh = torch.relu(vh @ p[136] + p[137])
    value = torch.tanh(vh @ p[138] + p[139]).reshape(-1)
    return policy, value, stats


def keras_loss(p, planes_nhwc, policy_index, value_target):
    """Returns (total, policy CE, value MSE, regulariser, BN statistics)."""
    policy, value, stats = forward_train(p, planes_nhwc)
    onehot = torch.zeros_like(policy)
    onehot[torch.arange(policy.shape[0]), policy_index] = 1.0
    q = policy / policy.sum(dim=-1, keepdim=True)
    q = torch.clamp(q, KERAS_EPS, 1.0 - KERAS_EPS)
    ce = (-(onehot * torch.log(q)).sum(dim=-1)).mean()
    mse = ((value - value_target) ** 2).mean()
    reg = L2 * sum((p[i] ** 2).sum() for i in kernel_indices([tuple(t.shape) for t in p]))
    return ce + mse + reg, ce, mse, reg, stats


def loss_and_grads(pack, planes_nhwc, policy_index, value_target, trainable):
    """pack: numpy arrays.  Returns (dict of float losses, {index: float64 gradient} for `trainable`, BN statistics)."""
    p = [torch.tensor(np.asarray(w), dtype=torch.float64) for w in pack]
    for i in trainable:
        p[i].requires_grad_(True)
    x = torch.as_tensor(np.asarray(planes_nhwc), dtype=torch.float64)
    total, ce, mse, reg, stats = keras_loss(p, x, torch.as_tensor(np.asarray(policy_index), dtype=torch.int64),
                                            torch.as_tensor(np.asarray(value_target), dtype=torch.float64))
    grads = torch.autograd.grad(total, [p[i] for i in trainable])
    return ({"loss": total.item(), "policy_loss": ce.item(), "value_loss": mse.item(), "reg": reg.item()},
            {i: g.numpy() for i, g in zip(trainable, grads)},
            [(m.numpy(), v.numpy()) for m, v in stats])


def keras_adam_step(w, g, m, v, t, lr=0.002, b1=0.9, b2=0.999, eps=1e-7):
    """One tf.keras Adam update in numpy float64: returns (w', m', v')."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return w - lr_t * m / (np.sqrt(v) + eps), m, v
