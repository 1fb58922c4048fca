"""Known-answer perft positions for the rule corners (shared by the oracle, host-build and GPU tests).

The LAST count of every row is the published total of the widely circulated "perft test positions" set for move
generators (en passant pins, castling through / into check, promotions in and out of check, stalemate traps); the
shallower counts were produced by the host build of the device code and the python-chess restatement, which agree
with each other and reproduce every published total (oracle run to full depth once, 147 s; the CPU suite runs it
to ORACLE_NODE_CAP nodes per position).  MAX_MOVES holds the two known 218-move positions, the most legal moves any
chess position has (sizes the per-node edge arena, selfplay.edge_slots_per_node)."""

EDGE = [
    ("3k4/3p4/8/K1P4r/8/8/8/8 b - - 0 1", [18, 92, 1670, 10138, 185429, 1134888]),        # ep illegal: rook pins both pawns on the rank
    ("8/8/4k3/8/2p5/8/B2P2K1/8 w - - 0 1", [13, 102, 1266, 10276, 135655, 1015133]),       # ep illegal: bishop pins the capturer
    ("8/8/1k6/2b5/2pP4/8/5K2/8 b - d3 0 1", [15, 126, 1928, 13931, 206379, 1440467]),      # ep capture gives check
    ("5k2/8/8/8/8/8/8/4K2R w K - 0 1", [15, 66, 1198, 6399, 120330, 661072]),              # short castling gives check
    ("3k4/8/8/8/8/8/8/R3K3 w Q - 0 1", [16, 71, 1286, 7418, 141077, 803711]),              # long castling gives check
    ("r3k2r/1b4bq/8/8/8/8/7B/R3K2R w KQkq - 0 1", [26, 1141, 27826, 1274206]),             # castling rights lost by rook capture
    ("r3k2r/8/3Q4/8/8/5q2/8/R3K2R b KQkq - 0 1", [44, 1494, 50509, 1720476]),              # castling prevented by attacked squares
    ("2K2r2/4P3/8/8/8/8/8/3k4 w - - 0 1", [11, 133, 1442, 19174, 266199, 3821001]),        # promote out of check
    ("8/8/1P2K3/8/2n5/1q6/8/5k2 b - - 0 1", [29, 165, 5160, 31961, 1004658]),              # discovered check
    ("4k3/1P6/8/8/8/8/K7/8 w - - 0 1", [9, 40, 472, 2661, 38983, 217342]),                 # promote to give check
    ("8/P1k5/K7/8/8/8/8/8 w - - 0 1", [6, 27, 273, 1329, 18135, 92683]),                   # underpromote to give check
    ("K1k5/8/P7/8/8/8/8/8 w - - 0 1", [2, 6, 13, 63, 382, 2217]),                          # self stalemate
    ("8/k1P5/8/1K6/8/8/8/8 w - - 0 1", [10, 25, 268, 926, 10857, 43261, 567584]),          # stalemate and checkmate
    ("8/8/2k5/5q2/5n2/8/5K2/8 b - - 0 1", [37, 183, 6559, 23527]),                         # stalemate and checkmate
]

MAX_MOVES = [
    "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1",
    "3Q4/1Q4Q1/4Q3/2Q4R/Q4Q2/3Q4/1Q4Rp/1K1BBNNk w - - 0 1",
]

ORACLE_NODE_CAP = 60_000      # pure-Python perft: ~60 k nodes/s
