"""Known answers for the python-chess 0.28.3 restatement that come from OUTSIDE this repository.

Two kinds, kept apart because their provenance differs:

README    listings printed in python-chess's own README / documentation ("Features" section and the SquareSet docs),
          which the 0.2x releases carry unchanged.  These are the only move-ORDER listing (start position), the
          Scholar's-mate walk-through with its Board(...) repr, the draw-rule flags after it, the attackers mask and the
          SquareSet iteration / mirror semantics netencoder.py:25-27 relies on.
SUITE     positions from the library's own test-suite (test.py, BoardTestCase: test_insufficient_material,
          test_fivefold_repetition, test_fifty_moves).  python-chess is not installable in this image, so these rows
          are RECALLED, not copied from a checked-out tree; every expected value was re-derived by hand from the FIDE
          rule it tests (square colours of the bishops, occurrence counts, clock values) and the derivation is in
          the comment.  They pin the rule corners Game.get_result depends on (game.py:92-109).

Version switches of the restatement (oracle/pychess_compat/chess: FIFTY_MOVE_CLAIM_LOOKAHEAD,
REPETITION_STOPS_ON_LEGAL_EP) stay at their 0.28.3 defaults for every row.
"""

# ---- README -----------------------------------------------------------------------------------------------------
README_START_LEGAL_MOVES_SAN = "Nh3, Nf3, Nc3, Na3, h3, g3, f3, e3, d3, c3, b3, a3, h4, g4, f4, e4, d4, c4, b4, a4"
README_START_LEGAL_MOVES = ("g1h3 g1f3 b1c3 b1a3 h2h3 g2g3 f2f3 e2e3 d2d3 c2c3 b2b3 a2a3 "
                            "h2h4 g2g4 f2f4 e2e4 d2d4 c2c4 b2b4 a2a4").split()
# "Scholar's mate" walk-through: push_san e4 e5 Qh5 Nc6 Bc4 Nf6 Qxf7 -> is_checkmate() True and the printed repr
README_SCHOLARS_MATE = ["e2e4", "e7e5", "d1h5", "b8c6", "f1c4", "g8f6", "h5f7"]
README_SCHOLARS_MATE_FEN = "r1bqkb1r/pppp1Qpp/2n2n2/4p3/2B1P3/8/PPPP1PPP/RNB1K1NR b KQkq - 0 4"
README_SCHOLARS_MATE_FLAGS = {           # the "Detects ..." bullets evaluated on that board
    "is_checkmate": True, "is_stalemate": False, "is_insufficient_material": False, "is_game_over": True,
    "can_claim_fifty_moves": False, "is_fivefold_repetition": False, "is_seventyfive_moves": False, "is_check": True,
}
README_SCHOLARS_MATE_HALFMOVE_CLOCK = 0
README_ATTACKED_E8_BY_WHITE = True                       # board.is_attacked_by(chess.WHITE, chess.E8)
README_ATTACKERS_OF_F3_BY_WHITE = 0x0000_0000_0000_4040  # board.attackers(chess.WHITE, chess.F3): g1 knight, g2 pawn
README_FEN_ROUND_TRIP = "8/8/8/2k5/4K3/8/8/8 w - - 4 45"  # piece_at(C5) is a black king
# SquareSet docs: SquareSet(BB_A8 | BB_RANK_1) has 9 members, iterates ascending, list(...) == [0..7, 56]
SQUARESET_DOC_MASK = (1 << 56) | 0xFF
SQUARESET_DOC_LIST = [0, 1, 2, 3, 4, 5, 6, 7, 56]

# ---- SUITE: insufficient material, (fen, white has insufficient material, black has) ------------------------------
# square colour: (file + rank) even = dark, odd = light, a1 = (0 + 0) dark
INSUFFICIENT = [
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", False, False),
    ("k1K1B1B1/8/8/8/8/8/8/8 w - - 7 32", True, True),      # e8 (4+7) and g8 (6+7): both light, black bare
    ("kbK1B1B1/8/8/8/8/8/8/8 w - - 7 32", False, False),    # + black bishop b8 (1+7) dark: opposite colours exist
    ("8/5k2/8/8/8/8/3K4/8 w - - 0 1", True, True),          # K v K
    ("8/3k4/8/8/2N5/8/3K4/8 b - - 0 1", True, True),        # KN v K
    ("8/4rk2/8/8/8/8/3K4/8 w - - 0 1", True, False),        # a rook always suffices for its owner
    ("8/4qk2/8/8/8/8/3K4/8 w - - 0 1", True, False),
    ("8/4bk2/8/8/8/8/3KB3/8 w - - 0 1", False, False),      # e7 (4+6) dark v e2 (4+1) light
    ("8/8/3Q4/2bK4/B7/8/1k6/8 w - - 1 68", False, False),   # queen; c5 dark v a4 light
    ("8/5k2/8/8/8/4B3/3K1B2/8 w - - 0 1", True, True),      # e3 (4+2), f2 (5+1): both dark
    ("5K2/8/8/1B6/8/k7/6b1/8 w - - 0 39", True, True),      # b5 (1+4), g2 (6+1): both light
    ("8/8/8/4k3/5b2/3K4/8/2B5 w - - 0 33", True, True),     # f4 (5+3), c1 (2+0): both dark
    ("3b4/8/8/6b1/8/8/R7/K1k5 w - - 0 1", False, True),     # white rook suffices; d8 (3+7), g5 (6+4) both dark and
                                                            # white has no bishop / knight / pawn to help black mate
]

# ---- SUITE: fivefold repetition --------------------------------------------------------------------------------
FIVEFOLD_FEN = "rnbq1rk1/ppp3pp/3bpn2/3p1p2/2PP4/2NBPN2/PP3PPP/R1BQK2R w KQ - 3 7"
FIVEFOLD_CYCLE = ["d3e2", "f6e4", "e2d3", "e4f6"]            # Be2 Ne4 Bd3 Nf6: back to the same position
# after 3 cycles the position has occurred 4 times: not over; after the 4th cycle 5 times: is_fivefold_repetition and
# is_game_over.  Then Qc2 Qd7 Qd2 Qe7 Qd1 (no repetition claim... the library's test continues) and Qd8 brings the
# SAME position back a sixth time although the occurrences are no longer consecutive: fivefold again.
FIVEFOLD_DETOUR = ["d1c2", "d8d7", "c2d2", "d7e7", "d2d1"]   # not a repetition at any point of the detour
FIVEFOLD_RETURN = "e7d8"

# ---- SUITE: fifty-move claim (positions of Timman - Lutz 1995), (fen, can_claim_fifty_moves, is_seventyfive_moves) --
FIFTY_MOVES = [
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", False, False),
    ("8/5R2/8/r2KB3/6k1/8/8/8 w - - 19 79", False, False),
    ("8/8/6r1/4B3/8/4K2k/5R2/8 b - - 68 103", False, False),
    ("6R1/7k/8/8/1r3B2/5K2/8/8 w - - 99 119", False, False),    # 0.28.3: the claim needs the clock AT 100 (no lookahead)
    ("8/7k/8/6R1/1r3B2/5K2/8/8 b - - 100 119", True, False),
    ("8/7k/8/1r3KR1/5B2/8/8/8 w - - 105 122", True, False),
    ("k7/8/NKB5/8/8/8/8/8 b - - 105 176", False, False),         # checkmated: too late to claim (no legal move)
    ("k7/3N4/1K6/1B6/8/8/8/8 b - - 99 1", False, False),          # stalemate: nothing to claim
]
