#!/usr/bin/env python
"""Reference-side converter between the reference's Keras checkpoints and chessrl_b200's weight pack.

Run it WHERE THE REFERENCE RUNS (TensorFlow >= 2.0 + the reference's src/chessrl on PYTHONPATH); it is not used by
the B200 package itself and cannot run in the build image (no TensorFlow there).

    # Keras checkpoint of the reference  ->  pack file chessrl_b200 loads (ChessModel.load_weights / Agent.load)
    python scripts/export_keras_weights.py model-3.h5 pack/model-3.h5

    # pack file written by chessrl_b200 (Agent.save)  ->  Keras checkpoint the reference loads
    python scripts/export_keras_weights.py --to-keras pack/model-4.h5 model-4.h5

The pack is an .npz archive (under whatever file name is given -- chessrl_b200 keeps the reference's model-<n>.h5
naming rule, selfplay.py:33-56) holding w000..w139 in the order chessrl_b200/model.py documents:
stem conv (kernel, bias); per residual block conv_a, BN_a, conv_b, BN_b; policy head conv, BN, Dense(1968);
value head conv, BN, Dense(256), Dense(1).  Kernels keep Keras layouts (HWIO / [in][out]), BatchNorm = gamma, beta,
moving_mean, moving_variance.  Layers are matched by kind and creation order (model.py:31-63, 111-122).
"""
import argparse
import sys

import numpy as np


def layer_groups(keras_model):
    from tensorflow.keras.layers import BatchNormalization, Conv2D, Dense
    convs = [l for l in keras_model.layers if isinstance(l, Conv2D)]
    bns = [l for l in keras_model.layers if isinstance(l, BatchNormalization)]
    dense_hidden = [l for l in keras_model.layers if isinstance(l, Dense) and l.name not in ("policy_out", "value_out")]
    assert len(convs) == 23 and len(bns) == 22 and len(dense_hidden) == 1, (len(convs), len(bns), len(dense_hidden))
    # model.layers is topologically sorted; inside one kind that equals creation order except for the two heads,
    # which are told apart by their filter count (policy: 2 filters, value: 1)
    trunk_convs = [l for l in convs if l.filters == 256]
    pol_conv = [l for l in convs if l.filters == 2][0]
    val_conv = [l for l in convs if l.filters == 1][0]
    trunk_bns = [l for l in bns if l.gamma.shape[0] == 256]
    pol_bn = [l for l in bns if l.gamma.shape[0] == 2][0]
    val_bn = [l for l in bns if l.gamma.shape[0] == 1][0]
    order = [trunk_convs[0]]
    for b in range(10):
        order += [trunk_convs[1 + 2 * b], trunk_bns[2 * b], trunk_convs[2 + 2 * b], trunk_bns[2 * b + 1]]
    order += [pol_conv, pol_bn, keras_model.get_layer("policy_out"),
              val_conv, val_bn, dense_hidden[0], keras_model.get_layer("value_out")]
    return order


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--to-keras", action="store_true")
    args = ap.parse_args()
    from model import ChessModel                      # the reference's src/chessrl/model.py
    if not args.to_keras:
        m = ChessModel(compile_model=False, weights=args.src)
        pack = []
        for layer in layer_groups(m.model):
            pack += [np.asarray(w, dtype=np.float32) for w in layer.get_weights()]
        assert len(pack) == 140, len(pack)
        with open(args.dst, "wb") as f:
            np.savez(f, **{"w%03d" % i: w for i, w in enumerate(pack)})
    else:
        m = ChessModel(compile_model=False)
        with np.load(args.src) as z:
            pack = [z["w%03d" % i] for i in range(140)]
        o = 0
        for layer in layer_groups(m.model):
            n = len(layer.get_weights())
            layer.set_weights(pack[o:o + n])
            o += n
        assert o == 140
        m.save_weights(args.dst)
    print("wrote", args.dst)


if __name__ == "__main__":
    sys.exit(main())
