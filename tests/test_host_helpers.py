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
