"""Seeded generator of positions that need NOT be reachable from the start position (shared by the host-build and
GPU rule tests): up to 27 pieces anywhere, promoted material, castling material at home with random rights,
double-pushed pawns with a capturer beside them and a raw en-passant square (python-chess keeps the square after
every double push; whether the capture is legal only matters for FEN / the transposition key)."""
import chessrl_oracle as O

chess = O.chess


def random_fen(rng):
    while True:
        grid = [None] * 64
        sqs = rng.sample(range(64), rng.randrange(3, 28))
        wk, bk = sqs[0], sqs[1]
        home = rng.random() < 0.35                        # castling material at home
        if home:
            wk, bk = 4, 60
        if max(abs((wk >> 3) - (bk >> 3)), abs((wk & 7) - (bk & 7))) <= 1:
            continue
        grid[wk], grid[bk] = "K", "k"
        if home:
            for s, c in ((0, "R"), (7, "R"), (56, "r"), (63, "r")):
                if rng.random() < 0.8:
                    grid[s] = c
        for s in sqs[2:]:
            if grid[s] is not None:
                continue
            pt = rng.choice("pppnbrq")
            if pt == "p" and (s >> 3) in (0, 7):
                pt = rng.choice("nbrq")
            grid[s] = pt.upper() if rng.random() < 0.5 else pt
        white = rng.random() < 0.5
        if rng.random() < 0.4:                            # a double-pushed pawn with a capturer beside it
            f = rng.randrange(8)
            s = (32 if white else 24) + f
            path = (s + 8, s + 16) if white else (s - 8, s - 16)
            nb = [s + d for d in (-1, 1) if 0 <= f + d < 8]
            if all(grid[x] in (None, "p", "P") for x in (s,) + path):
                grid[s] = "p" if white else "P"
                grid[path[0]] = grid[path[1]] = None
                x = rng.choice(nb)
                if grid[x] not in ("K", "k"):
                    grid[x] = "P" if white else "p"
        cr = "".join(c for c, k, r, kc, rc in (("K", 4, 7, "K", "R"), ("Q", 4, 0, "K", "R"),
                                               ("k", 60, 63, "k", "r"), ("q", 60, 56, "k", "r"))
                     if grid[k] == kc and grid[r] == rc and rng.random() < 0.7) or "-"
        cands = []
        for s in range(64):
            if white and grid[s] == "p" and (s >> 3) == 4 and grid[s + 8] is None and grid[s + 16] is None:
                cands.append(s + 8)
            if not white and grid[s] == "P" and (s >> 3) == 3 and grid[s - 8] is None and grid[s - 16] is None:
                cands.append(s - 8)
        ep = "-"
        if cands and rng.random() < 0.6:
            e = rng.choice(cands)
            ep = "abcdefgh"[e & 7] + str((e >> 3) + 1)
        rows = []
        for r in range(7, -1, -1):
            row, gap = "", 0
            for f in range(8):
                c = grid[r * 8 + f]
                if c is None:
                    gap += 1
                else:
                    row += (str(gap) if gap else "") + c
                    gap = 0
            rows.append(row + (str(gap) if gap else ""))
        fen = "%s %s %s %s %d %d" % ("/".join(rows), "w" if white else "b", cr, ep, rng.randrange(0, 60),
                                     rng.randrange(1, 80))
        b = chess.Board(fen)
        b.turn = not b.turn                               # the side that just moved must not be in check
        bad = b.is_check()
        b.turn = not b.turn
        if not bad:
            return fen, b
