"""Host-side helpers that need no device: model-file discovery (selfplay.py:33-56 incl. its string-max quirk),
game packing for the rank-0 gather, move-list packing for crl_games_set_host, record <-> FEN."""
import numpy as np

from chessrl_b200 import boards as B
from chessrl_b200 import sharding
from chessrl_b200.engine import Engine
from chessrl_b200.selfplay import get_model_path


def test_get_model_path_follows_the_reference_rule(tmp_path):
    d = str(tmp_path)
    assert get_model_path(d) == d + "/model-0.h5"                 # nothing there yet: the default name
    for name in ("model-3.h5", "model-12.h5", "notes.txt"):
        (tmp_path / name).write_text("")
    # max() over the version STRINGS, as the reference does: "3.h5" > "12.h5"
    assert get_model_path(d) == d + "/model-3.h5"


def test_pack_games_layout():
    words = [[1, 2, 3], [], [65534, 7]]
    mv, ln, res, col = sharding.pack_games(words, [1, None, -1], [True, False, True])
    assert mv.shape == (3, 3) and mv.dtype == np.int16 and list(ln) == [3, 0, 2]
    assert list(mv[0]) == [1, 2, 3] and list(mv[1]) == [-1, -1, -1]
    assert list(mv[2, :2].view(np.uint16)) == [65534, 7]
    assert list(res) == [1, B.RESULT_NONE, -1] and list(col) == [1, 0, 1]
    mv, ln, res, col = sharding.pack_games([], [], [])            # a rank with no games still packs
    assert mv.shape == (0, 1) and ln.shape == (0,)


def test_pack_move_lists():
    mv, cnt = Engine.pack_move_lists([[5, 6], [], [7, 8, 9]])
    assert mv.shape == (3, 3) and mv.dtype == np.uint16 and list(cnt) == [2, 0, 3]
    assert list(mv[0]) == [5, 6, B.MOVE_NONE] and list(mv[1]) == [B.MOVE_NONE] * 3
    mv, cnt = Engine.pack_move_lists([[], []])
    assert mv.shape == (2, 1) and list(cnt) == [0, 0]


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_fen_record_roundtrip_keeps_every_field():
    for fen in ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
                "rnbqkbnr/pppp1ppp/8/4p3/4P3/8/PPPP1PPP/RNBQKBNR w KQkq e6 0 2",
                "8/8/8/K2pP2r/8/8/8/4k3 w - d6 37 91", "4k3/8/8/8/8/8/8/R3K2R b Q - 99 150"):
        rec = B.record_from_fen(fen)
        f = B.meta_fields(rec[8])
        parts = fen.split()
        assert f["turn"] == (parts[1] == "w") and f["halfmove"] == int(parts[4]) and f["fullmove"] == int(parts[5])
        assert B.board_fen_from_record(rec) == parts[0]
        assert B.fen_from_record(rec, True) == fen                 # ep square printed when a capture is legal
        if parts[3] != "-":
            assert B.fen_from_record(rec, False) == " ".join(parts[:3] + ["-"] + parts[4:])


def test_model_serial_is_bumped_by_every_weight_assignment(tmp_path):
    """runtime.ensure_weights compares ChessModel.serial (process-wide, never reused) instead of id(model)."""
    from chessrl_b200.model import ChessModel
    a, b = ChessModel(seed=0), ChessModel(seed=1)
    assert a.serial != b.serial
    s = a.serial
    a.weights = [w.copy() for w in a.weights]
    assert a.serial > s
    path = str(tmp_path / "model-0.h5")
    a.save_weights(path)
    s = b.serial
    b.load_weights(path)                                  # direct load_weights also counts as a new set of weights
    assert b.serial > s and all(np.array_equal(x, y) for x, y in zip(a.weights, b.weights))
    serials = {ChessModel(seed=0).serial for _ in range(5)}      # objects freed in between still get fresh serials
    assert len(serials) == 5


def test_keras_checkpoint_layer_mapping_and_hdf5_detection(tmp_path):
    """model.py:77-81 reads Keras .h5 files: layers are matched by kind and creation order whatever numbers Keras
    appended to their names; a real HDF5 file without h5py gives a clear error, not a pickle crash."""
    import pytest
    from chessrl_b200 import keras_h5
    from chessrl_b200.model import ChessModel, random_pack
    pack = random_pack(seed=9, perturb_bn=True)
    for first in (0, 23, 46):                             # first, second, third model built in the writing process
        layers = keras_h5.keras_layers_from_pack(pack, first_suffix=first)
        assert len(layers) == 23 + 22 + 3
        if first == 0:
            assert "conv2d" in layers and "batch_normalization_21" in layers and "dense" in layers
        # Keras also lists weightless layers (Input, Activation, Add, Flatten) with empty weight lists
        layers["activation_%d" % (first + 1)] = {}
        got = keras_h5.pack_from_keras_layers(dict(reversed(list(layers.items()))))     # file order does not matter
        assert all(np.array_equal(a, b) for a, b in zip(got, pack))
    broken = keras_h5.keras_layers_from_pack(pack)
    del broken["conv2d_7"]
    with pytest.raises(ValueError):
        keras_h5.pack_from_keras_layers(broken)
    swapped = keras_h5.keras_layers_from_pack(pack)
    swapped["conv2d_21"], swapped["conv2d_22"] = swapped["conv2d_22"], swapped["conv2d_21"]
    with pytest.raises(ValueError):                       # policy / value head convolutions swapped: shapes give it away
        keras_h5.pack_from_keras_layers(swapped)
    fake = tmp_path / "model-1.h5"
    fake.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    try:
        import h5py  # noqa: F401
        has_h5py = True
    except ImportError:
        has_h5py = False
    if not has_h5py:
        with pytest.raises(RuntimeError, match="export_keras_weights"):
            ChessModel(weights=str(fake))
    with pytest.raises(OSError):
        ChessModel(weights=str(tmp_path / "missing.h5"))   # supervised.py:57-59 catches OSError for a missing file


def test_default_lane_count_balances_drain_and_tower_rounds():
    from chessrl_b200.selfplay import TOWER_ROUND, default_lanes
    assert default_lanes(1) == 1 and default_lanes(500) == 500 and default_lanes(592) == 592
    assert default_lanes(593) == 592 and default_lanes(4096) == 592           # 1,024 -> one whole round
    assert default_lanes(8192) == 3 * 592 and default_lanes(65536) == 7 * 592
    assert all(default_lanes(n) % TOWER_ROUND == 0 for n in range(600, 70000, 997))
