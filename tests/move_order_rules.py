"""Structural statement of python-chess 0.28.3's legal-move ORDER (SURVEY.md 8c "move order"), written as a VALIDATOR
that is independent of every generator in this repository: it classifies the moves of a list from the position's
bitboards and checks the ordering invariants, without generating a single move itself.

Not in check -- classes in this order (a class may be empty):
  1 officer moves (knight, bishop, rook, queen, king steps): from-square descending (h8 -> a1), then to-square descending
  2 castling (king side before queen side, written as the king's move)
  3 pawn captures: from descending, to descending, a promotion expands to q, r, b, n in that order
  4 single pawn pushes by to-square descending (promotions q, r, b, n)
  5 double pawn pushes by to-square descending
  6 en passant captures, capturer's from-square descending
In check: king moves first (to descending), then classes 1 (non-king), 3, 4, 5 of the blockers / capturers, then 6.
"""

PROMO_RANK = {4: 0, 3: 1, 2: 2, 1: 3}          # move word promo code (4 Q, 3 R, 2 B, 1 N) -> position in q, r, b, n


def classify(rec, mv):
    """rec: 9 ints (board record), mv: move word.  -> (class, from, to, promo rank or -1, is_king)"""
    f, t, p = mv & 63, (mv >> 6) & 63, (mv >> 12) & 7
    pawns, kings = int(rec[0]), int(rec[5])
    occ = int(rec[6]) | int(rec[7])
    fb, tb = 1 << f, 1 << t
    pr = PROMO_RANK[p] if p else -1
    if pawns & fb:
        if (f & 7) != (t & 7):
            return (3 if occ & tb else 6), f, t, pr, False
        return (5 if abs(t - f) == 16 else 4), f, t, pr, False
    if kings & fb and abs((t & 7) - (f & 7)) == 2:
        return 2, f, t, pr, True
    return 1, f, t, pr, bool(kings & fb)


def violations(rec, moves, in_check):
    """List of human-readable violations of the order rules (empty = the list is well ordered)."""
    bad = []
    cls = [classify(rec, int(m)) for m in moves]
    if len(set(int(m) for m in moves)) != len(moves):
        bad.append("duplicate move")
    if in_check:
        # king moves lead; everything after them follows the normal class order
        k = 0
        while k < len(cls) and cls[k][4] and cls[k][0] == 1:
            k += 1
        if any(c[4] for c in cls[k:]):
            bad.append("king move after a non-king move while in check")
        if any(c[0] == 2 for c in cls):
            bad.append("castling while in check")
        kings = cls[:k]
        if any(a[2] <= b[2] for a, b in zip(kings, kings[1:])):
            bad.append("king evasions not by to-square descending")
        cls = cls[k:]
    order = [c[0] for c in cls]
    if order != sorted(order):
        bad.append("classes out of order: %s" % order)
    for c in (1, 2, 3, 4, 5, 6):
        seq = [x for x in cls if x[0] == c]
        if c == 1 or c == 3:
            keys = [(-x[1], -x[2], x[3]) for x in seq]
        elif c == 2:
            keys = [(-x[2],) for x in seq]                     # g-file target before c-file target
        elif c == 4 or c == 5:
            keys = [(-x[2], x[3]) for x in seq]
        else:
            keys = [(-x[1],) for x in seq]
        if any(a >= b for a, b in zip(keys, keys[1:])):
            bad.append("class %d not in order" % c)
    for x in cls:
        if x[3] >= 0 and x[0] not in (3, 4):
            bad.append("promotion outside the pawn classes")
    # a promotion always comes as the complete q, r, b, n group
    promos = [x for x in cls if x[3] >= 0]
    if len(promos) % 4 or any([y[3] for y in promos[i:i + 4]] != [0, 1, 2, 3] or len({(y[1], y[2]) for y in promos[i:i + 4]}) != 1
                              for i in range(0, len(promos), 4)):
        bad.append("promotion group is not q, r, b, n")
    return bad
