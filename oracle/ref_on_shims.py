"""ORACLE / TEST INFRASTRUCTURE ONLY.

Loads the reference's own, UNMODIFIED `game.py`, `player.py`, `mctree.py`, `netencoder.py`, `dataset.py`
and `agentdistributed.py` from $CHESSRL_REF (default /root/reference/src/chessrl) on top of the
python-chess restatement in oracle/pychess_compat, with import-time stubs for the packages those files
import but the hot path never calls (chess.svg, cairosvg, PIL, matplotlib: game.py:1-8;
tensorflow.keras.utils: netencoder.py:10).  Nothing is copied: the files are imported where they lie.

Only usable where the reference tree is mounted (this container).  The GPU box never has it, so
tests/golden/make_golden.py uses this loader once to write small golden fixtures that travel.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

DEFAULT_REF = "/root/reference/src/chessrl"
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_MODULES = ("game", "player", "mctree", "netencoder", "dataset", "agentdistributed")


def reference_dir():
    d = os.environ.get("CHESSRL_REF", DEFAULT_REF)
    return d if os.path.isfile(os.path.join(d, "mctree.py")) else None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _to_categorical(y, num_classes=None):
    out = np.zeros(num_classes, dtype=np.float32)
    out[int(y)] = 1.0
    return out


def load_reference():
    """Returns a namespace with the reference modules, or None when the reference tree is absent."""
    ref = reference_dir()
    if ref is None:
        return None
    compat = os.path.join(_HERE, "pychess_compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    import chess  # the restatement

    injected = {}

    def inject(name, mod):
        if name not in sys.modules:
            try:
                if name.split(".")[0] in ("tensorflow",):
                    raise ImportError
                importlib.import_module(name)
                return
            except Exception:
                pass
            sys.modules[name] = mod
            injected[name] = mod

    pil = _stub("PIL")
    pil.Image = _stub("PIL.Image")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    tf = _stub("tensorflow")
    tf.keras = _stub("tensorflow.keras")
    tf.keras.utils = _stub("tensorflow.keras.utils", Sequence=object, to_categorical=_to_categorical)
    inject("cairosvg", _stub("cairosvg"))
    inject("PIL", pil)
    inject("PIL.Image", pil.Image)
    inject("matplotlib", mpl)
    inject("matplotlib.pyplot", mpl.pyplot)
    inject("tensorflow", tf)
    inject("tensorflow.keras", tf.keras)
    inject("tensorflow.keras.utils", tf.keras.utils)

    saved = {n: sys.modules.pop(n) for n in _REF_MODULES if n in sys.modules}
    sys.path.insert(0, ref)
    try:
        mods = {n: importlib.import_module(n) for n in _REF_MODULES}
    finally:
        sys.path.remove(ref)
        for n in _REF_MODULES:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
        for n in injected:           # do not leave a fake tensorflow/PIL behind for torch & friends
            sys.modules.pop(n, None)
    ns = types.SimpleNamespace(**mods)
    ns.chess = chess
    ns.ref_dir = ref
    return ns


def make_ref_agent(ref, evaluate, color=True):
    """The reference's AgentDistributed with only its socket hop replaced by `evaluate(game)`;
    predict_policy / best_move / predict_outcome stay the reference's own code
    (agentdistributed.py:39-99)."""

    class _InProcessAgent(ref.agentdistributed.AgentDistributed):
        n_evals = 0

        def _AgentDistributed__send_game(self, game):          # replaces agentdistributed.py:90-99
            type(self).n_evals += 1
            p, v = evaluate(game)
            return p, float(v)                                  # predict_worker.py:111

        def get_copy(self):
            return self

        def connect(self):
            pass

        def disconnect(self):
            pass

    return _InProcessAgent(color, endpoint=None, num_threads=1)
