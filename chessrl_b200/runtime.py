"""Process-wide engine handles for the single-object API (Game / Agent / SelfPlayTree used one game at a time,
the way the reference's scripts use them).  One engine per (process, GPU); the lockstep driver creates its own
wide engines.  Creation fails loudly without CUDA -- there is no CPU path."""

from __future__ import annotations

_scalar = None
_loaded_token = None


def scalar_engine(min_nodes=1024, min_inflight=1):
    """The one-lane engine behind Game / netencoder / Agent / SelfPlayTree (grown on demand)."""
    global _scalar, _loaded_token
    from .engine import Engine
    if _scalar is None or _scalar.max_nodes < min_nodes or _scalar.max_inflight < min_inflight:
        nodes, inflight = max(1024, int(min_nodes)), max(8, int(min_inflight))
        if _scalar is not None:
            nodes, inflight = max(nodes, _scalar.max_nodes), max(inflight, _scalar.max_inflight)
            _scalar.close()
        _scalar = Engine(max_games=1, max_nodes=nodes, avg_moves=218, max_inflight=inflight)   # 218 = most legal moves of any position
        _loaded_token = None
    return _scalar


def ensure_weights(engine, model):
    """Uploads `model`'s weight pack to the scalar engine unless it is already there."""
    global _loaded_token
    token = model.serial        # process-wide, bumped by every weight assignment (model.ChessModel.weights)
    if engine is _scalar:
        if _loaded_token != token:
            engine.load_weights(model.weights)
            _loaded_token = token
    else:
        engine.load_weights(model.weights)


def shutdown():
    global _scalar, _loaded_token
    if _scalar is not None:
        _scalar.close()
    _scalar, _loaded_token = None, None
