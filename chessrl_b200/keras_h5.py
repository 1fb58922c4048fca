"""Reading the reference's real checkpoints: Keras `Model.save_weights('model-<n>.h5')` files (model.py:77-84).

A Keras weights file stores, per layer, the arrays of `layer.weights` under auto-generated layer names
(`conv2d_7`, `batch_normalization_3`, `dense`, `policy_out`, `value_out`); the numeric suffixes depend on how many
layers the writing process had created before, so layers are matched by KIND and ORDER of creation (model.py:31-63,
111-122), never by absolute name, and every array is checked against the weight pack's shapes (model.pack_shapes).

`pack_from_keras_layers` is the pure mapping (tested without h5py); `load_keras_h5` needs h5py, which this image
does not ship -- without it the error says how to convert the file on the reference side
(scripts/export_keras_weights.py, which runs where TensorFlow is installed).
"""

from __future__ import annotations

import re

import numpy as np

from .model import N_BLOCKS, N_TENSORS, pack_shapes

_BN_ORDER = ("gamma", "beta", "moving_mean", "moving_variance")


def _suffix(name):
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def _short(weight_name):
    """'conv2d_3/kernel:0' -> 'kernel'"""
    return weight_name.split("/")[-1].split(":")[0]


def pack_from_keras_layers(layers):
    """layers: {layer_name: {weight_name: array}} as stored in the file (layers without weights may be absent).
    Returns the 140-tensor pack in chessrl_b200.model order."""
    named = {ln: {_short(wn): np.asarray(a, dtype=np.float32) for wn, a in ws.items()} for ln, ws in layers.items() if ws}
    convs = sorted((n for n in named if n.startswith("conv2d")), key=_suffix)
    bns = sorted((n for n in named if n.startswith("batch_normalization")), key=_suffix)
    denses = sorted((n for n in named if n.startswith("dense")), key=_suffix)
    n_conv, n_bn = 1 + 2 * N_BLOCKS + 2, 2 * N_BLOCKS + 2
    if len(convs) != n_conv or len(bns) != n_bn or len(denses) != 1 or "policy_out" not in named or "value_out" not in named:
        raise ValueError("not a ChessModel weights file: %d conv / %d batch-norm / %d dense layers, policy_out %s, "
                         "value_out %s (expected %d / %d / 1)" % (len(convs), len(bns), len(denses), "policy_out" in named,
                                                                   "value_out" in named, n_conv, n_bn))

    def conv(i):
        return [named[convs[i]]["kernel"], named[convs[i]]["bias"]]

    def bn(i):
        return [named[bns[i]][k] for k in _BN_ORDER]

    def dense(name):
        return [named[name]["kernel"], named[name]["bias"]]

    pack = conv(0)
    for blk in range(N_BLOCKS):                       # __res_block: conv, BN, conv, BN (model.py:111-122)
        pack += conv(1 + 2 * blk) + bn(2 * blk) + conv(2 + 2 * blk) + bn(2 * blk + 1)
    pack += conv(n_conv - 2) + bn(n_bn - 2) + dense("policy_out")                      # policy head (model.py:39-47)
    pack += conv(n_conv - 1) + bn(n_bn - 1) + dense(denses[0]) + dense("value_out")    # value head (model.py:50-61)
    assert len(pack) == N_TENSORS
    for i, (a, sh) in enumerate(zip(pack, pack_shapes())):
        if tuple(a.shape) != tuple(sh):
            raise ValueError("tensor %d of the Keras file has shape %s, the ChessModel graph needs %s" % (i, a.shape, sh))
    return [np.ascontiguousarray(a) for a in pack]


def keras_layers_from_pack(pack, first_suffix=0):
    """Inverse of pack_from_keras_layers: the {layer: {weight: array}} dict Keras would write for a model whose
    auto-numbered layer names start at `first_suffix` (tests; also the layout scripts/export_keras_weights.py fills)."""
    def nm(base, k):
        k += first_suffix
        return base if k == 0 else "%s_%d" % (base, k)

    layers = {}
    ci = bi = 0

    def put_conv(o):
        nonlocal ci
        n = nm("conv2d", ci)
        layers[n] = {n + "/kernel:0": pack[o], n + "/bias:0": pack[o + 1]}
        ci += 1

    def put_bn(o):
        nonlocal bi
        n = nm("batch_normalization", bi)
        layers[n] = {"%s/%s:0" % (n, k): pack[o + j] for j, k in enumerate(_BN_ORDER)}
        bi += 1

    put_conv(0)
    for blk in range(N_BLOCKS):
        o = 2 + 12 * blk
        put_conv(o), put_bn(o + 2), put_conv(o + 6), put_bn(o + 8)
    put_conv(122), put_bn(124)
    layers["policy_out"] = {"policy_out/kernel:0": pack[128], "policy_out/bias:0": pack[129]}
    put_conv(130), put_bn(132)
    d = nm("dense", 0)
    layers[d] = {d + "/kernel:0": pack[136], d + "/bias:0": pack[137]}
    layers["value_out"] = {"value_out/kernel:0": pack[138], "value_out/bias:0": pack[139]}
    return layers


def load_keras_h5(path):
    try:
        import h5py
    except ImportError as exc:
        raise RuntimeError(
            "%s is a Keras HDF5 checkpoint of the reference; reading it needs h5py, which is not installed here. "
            "Convert it where the reference runs:  python scripts/export_keras_weights.py %s out.h5  "
            "(writes the 140-tensor pack this package loads; keep the .h5 file name)" % (path, path)) from exc
    layers = {}
    with h5py.File(path, "r") as f:
        root = f["model_weights"] if "model_weights" in f else f          # model.save() vs model.save_weights()
        for ln in root.attrs["layer_names"]:
            ln = ln.decode() if isinstance(ln, bytes) else str(ln)
            g = root[ln]
            ws = {}
            for wn in g.attrs["weight_names"]:
                wn = wn.decode() if isinstance(wn, bytes) else str(wn)
                ws[wn] = np.asarray(g[wn])
            layers[ln] = ws
    return pack_from_keras_layers(layers)
