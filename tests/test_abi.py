"""The C-ABI boundary without a GPU: the library builds and loads, exports every symbol the header declares,
its host-side format table is right, and it refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import chessrl_oracle as O
from chessrl_b200 import _lib
from chessrl_b200 import boards as B
from conftest import ROOT, has_gpu


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "chessrl_b200.h")).read()
    declared = set(re.findall(r"\b(crl_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_label_table_matches_reference_labels():
    lib = _lib.load()
    t = np.zeros(5 * 4096, dtype=np.int16)
    assert lib.crl_label_table_host(None, t.ctypes.data_as(_lib.c_i16p)) == 0
    t = t.reshape(5, 64, 64)
    assert (t >= 0).sum() == 1968
    for i, u in enumerate(O.uci_labels()):
        m = B.uci_to_move(u)
        assert t[(m >> 12) & 7, m & 63, (m >> 6) & 63] == i


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.crl_create(ctypes.byref(h), 0, 1, 8, 0, None)
    assert rc == _lib.CRL_ECUDA and b"no CPU path" in lib.crl_last_error()
    from chessrl_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()


def test_record_format_roundtrip():
    fen = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 3 17"
    rec = B.record_from_fen(fen)
    assert B.fen_from_record(rec) == fen
    m = B.meta_fields(rec[8])
    assert m["turn"] and m["castle"] == 15 and m["ep"] == -1 and m["halfmove"] == 3 and m["fullmove"] == 17
    assert B.move_to_uci(B.uci_to_move("e7e8q")) == "e7e8q" and B.move_to_uci(B.uci_to_move("g1f3")) == "g1f3"
    assert B.uci_to_move("00000") == B.MOVE_NONE and B.uci_to_move("e2e2") == B.MOVE_NONE
    # castling rights are cleaned like Board.clean_castling_rights()
    assert B.meta_fields(B.record_from_fen("4k3/8/8/8/8/8/8/4K2R w KQkq - 0 1")[8])["castle"] == 1


def test_create_rejects_bad_arguments_before_touching_a_device():
    """Argument validation of crl_create_ex comes first, so it is checkable without a GPU: negative status,
    crl_last_error text, *out untouched."""
    import ctypes
    lib = _lib.load()
    h = _lib.vp()
    for games, nodes, avg, inflight in ((0, 16, 64, 1), (4, 0, 64, 1), (4, 16, 64, 0), (4, 16, 64, 10 ** 6),
                                        (4, 16, 257, 1), (4, 1 << 24, 218, 1), (1 << 20, 16, 64, 32)):
        assert lib.crl_create_ex(ctypes.byref(h), 0, games, nodes, avg, inflight, None) == -1, (games, nodes, avg, inflight)
        assert b"bad arguments" in lib.crl_last_error() and not h.value
    assert lib.crl_create_ex(None, 0, 4, 16, 64, 1, None) == -1
