"""GPU parity, encoding: planes (netencoder.get_game_state), policy indices (uci_dict) and the deterministic
test evaluator -- bit-exact against the reference-generated goldens and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import chessrl_oracle as O
from chessrl_b200 import boards as B

pytestmark = pytest.mark.gpu
chess = O.chess


def _records_along(fen, moves):
    """Board records after every ply (raw ep square like python-chess)."""
    b = chess.Board(fen) if fen else chess.Board()
    out = []

    def rec():
        r = B.record_from_fen(b.fen())
        f = B.meta_fields(r[8])
        r[8] = B.pack_meta(f["turn"], f["castle"], -1 if b.ep_square is None else b.ep_square, f["halfmove"],
                           f["fullmove"], len(b.move_stack))
        return r
    out.append(rec())
    for m in moves:
        b.push(chess.Move.from_uci(m))
        out.append(rec())
    return out


def test_planes_golden_stateless(engine1, golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "planes.json")))["cases"]
    packed = np.load(os.path.join(golden_dir, "planes.npz"))["packed"]
    cases = [(c, bits) for c, bits in zip(meta, packed) if not c["flipped"]]
    n = len(cases)
    boards = np.zeros((n, 9), dtype=np.uint64)
    hist = np.zeros((8, 8, n), dtype=np.uint64)
    hlen = np.zeros(n, dtype=np.uint8)
    for i, (c, _) in enumerate(cases):
        recs = _records_along(c.get("fen"), c["moves"])
        boards[i] = recs[-1]
        prev = recs[-2::-1][:8]
        hlen[i] = len(prev)
        for j, r in enumerate(prev):
            hist[j, :, i] = r[:8]
    bt = engine1.boards_to_device(boards)
    ht = torch.from_numpy(hist.view(np.int64)).to(engine1.device)
    lt = torch.from_numpy(hlen).to(engine1.device)
    planes = engine1.encode(bt, ht, lt).float().cpu().numpy()
    assert planes.shape == (n, 8, 8, 128) and not planes[..., 127].any()
    for i, (c, bits) in enumerate(cases):
        want = np.unpackbits(bits)[:8 * 8 * 127].reshape(8, 8, 127)
        assert (planes[i, :, :, :127] == want).all(), c["name"]


def test_policy_index_matches_labels(engine1):
    idx = O.label_index()
    tab = engine1.label_table()
    assert (tab >= 0).sum() == 1968
    fens = [B.STARTING_FEN, "8/P6k/8/8/8/8/p7/K7 w - - 0 1", "8/P6k/8/8/8/8/p7/K7 b - - 0 1",
            "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"]
    bt = engine1.boards_to_device(np.stack([B.record_from_fen(f) for f in fens]))
    mv, cn, _ = engine1.movegen(bt)
    pi = engine1.policy_index(mv, cn).cpu().numpy()
    mvh = mv.cpu().numpy().view(np.uint16)
    for i, f in enumerate(fens):
        legal = [m.uci() for m in chess.Board(f).generate_legal_moves()]
        assert [int(x) for x in pi[i, :len(legal)]] == [idx[m] for m in legal]
        assert (pi[i, len(legal):] == -1).all()


def test_hash_evaluator_bit_exact(engine1):
    fens = [B.STARTING_FEN, "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
            "8/8/8/K2pP2r/8/8/8/4k3 w - d6 0 2"]
    for seed, bits in ((1, 24), (2, 3), (77, 11)):
        ev = O.hash_evaluator(seed, bits)
        bt = engine1.boards_to_device(np.stack([B.record_from_fen(f) for f in fens]))
        p, v = engine1.hash_eval(bt, seed, bits)
        p, v = p.cpu().numpy(), v.cpu().numpy()
        for i, f in enumerate(fens):
            wp, wv = ev(O.OGame(board=chess.Board(f)))
            assert (p[i] == wp).all() and v[i] == wv
