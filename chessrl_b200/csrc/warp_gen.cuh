// warp_gen.cuh -- the legal move generator of chess_core.cuh, re-cut so that the 32 lanes of a WARP cooperate on ONE board.
//
// Why: generate_legal() is ~3,000 dependent instructions for one thread -- about 30 us of LATENCY on a B200 SM however
// many boards run next to it.  Where there are far fewer boards than lanes (one tree expansion per game and simulation;
// the first plies of a perft: 1 -> 20 -> 400 boards) that latency is the whole cost.  Here every lane owns two squares
// (63 - lane and 31 - lane) and produces the moves that python-chess attributes to them:
//     class 1  officer moves FROM its squares            class 3  pawn captures FROM its squares
//     class 4  single pushes TO its squares              class 5  double pushes TO its squares
// python-chess emits each class in descending square order, i.e. in LANE order, upper half-board first, so
// an 8-field packed prefix sum over the lanes (one 64-bit add per shuffle step) gives every lane its output offsets and
// the list comes out in exactly the order generate_legal() produces.  King evasions (first, when in check), castling and
// en passant (at most two moves each) are written by lane 0.
//
// The per-lane pieces are CRL_HD so the TEST-ONLY host build (tests/hostsim) runs them lane by lane against the scalar
// generator; only the shuffle glue at the bottom is device code.
#pragma once
#include "chess_core.cuh"

namespace crl {

struct WgCommon {          // identical in every lane: everything generate_legal_side computes once per position
  u64 us, them, occ, pawns, officers, pinned, target, king_moves, single, dbl, promo_rank;
  int ksq, white, in_check, double_check, checker_sq, n_castle, valid;
  bool castle_k, castle_q;   // legal castling moves (king side is emitted first)
};

struct WgLane {            // the moves one lane emits; h = 0: square 63 - lane, h = 1: square 31 - lane
  u64 t1[2];               // officer targets from the lane's squares (already masked)
  u64 t3[2];               // pawn capture targets from the lane's squares
  u64 packed;              // counts [A1, B1, A3, B3, A4, B4, A5, B5], 8 bits each (A = upper square, B = lower)
};

CRL_HD void wg_common(const Board& b, WgCommon& c) {
  const int white = meta_turn(b.meta);
  c.white = white;
  c.us = b.bb[white ? OCC_W : OCC_B];
  c.them = b.bb[white ? OCC_B : OCC_W];
  c.occ = c.us | c.them;
  c.in_check = c.double_check = 0;
  c.checker_sq = -1;
  c.n_castle = 0;
  c.castle_k = c.castle_q = false;
  c.officers = c.pawns = c.pinned = c.king_moves = c.single = c.dbl = 0;
  c.target = ~0ULL;
  c.promo_rank = white ? RANK_8 : RANK_1;
  const u64 kings = b.bb[KING] & c.us;
  c.valid = kings != 0;
  if (!kings) return;
  const int ksq = msb64(kings);
  c.ksq = ksq;
  const u64 kbit = bit(ksq);
  const int base = white ? 0 : 56;
  const u64 occ = c.occ, us = c.us, them = c.them;

  const u64 king_ring = king_attacks_sq(ksq);
  u64 checkers = ((knight_attacks_sq(ksq) & b.bb[KNIGHT]) | (pawn_attacks_set(kbit, white) & b.bb[PAWN]) |
                  (king_ring & b.bb[KING])) & them;
  u64 pinned = 0;
  {
    const u64 rq = (b.bb[ROOK] | b.bb[QUEEN]) & them, bq = (b.bb[BISHOP] | b.bb[QUEEN]) & them;
    u64 snipers = ((rank_mask(ksq) | file_mask(ksq)) & rq) | ((diag_mask(ksq) | anti_mask(ksq)) & bq);
    while (snipers) {
      const int s = pop_msb(snipers);
      const u64 mid = between(ksq, s) & occ;
      if (!mid) checkers |= bit(s);
      else if (!(mid & (mid - 1))) pinned |= mid;
    }
    pinned &= us;
  }
  c.pinned = pinned;
  c.in_check = checkers != 0;

  const u64 king_targets = king_ring & ~us;
  const int rights = meta_castle(b.meta) >> (white ? 0 : 2);
  bool castle_k = false, castle_q = false;
  if (!checkers && ksq == base + 4) {
    castle_k = (rights & 1) && !(occ & (0x60ULL << base));
    castle_q = (rights & 2) && !(occ & (0x0EULL << base));
  }
  u64 danger = 0;
  if (king_targets || castle_k || castle_q) danger = attack_map(b, occ ^ kbit, !white);
  c.king_moves = king_targets & ~danger;
  c.castle_k = castle_k && !(danger & (0x60ULL << base));
  c.castle_q = castle_q && !(danger & (0x0CULL << base));
  c.n_castle = (int)c.castle_k + (int)c.castle_q;

  if (checkers) {
    if (checkers & (checkers - 1)) {
      c.double_check = 1;
      c.target = 0;
      return;                                    // king evasions only
    }
    c.checker_sq = msb64(checkers);
    c.target = between(ksq, c.checker_sq) | checkers;
  }
  c.officers = us & ~b.bb[PAWN];
  if (checkers) c.officers &= ~kbit;
  const u64 pawns = b.bb[PAWN] & us;
  c.pawns = pawns;
  if (pawns) {
    u64 single, dbl;
    if (white) {
      single = (pawns << 8) & ~occ;
      dbl = (single << 8) & ~occ & (0xFFULL << 24);
    } else {
      single = (pawns >> 8) & ~occ;
      dbl = (single >> 8) & ~occ & (0xFFULL << 32);
    }
    single &= c.target;
    dbl &= c.target;
    if (pawns & pinned) {
      const u64 ok_src = pawns & ~(pinned & ~file_mask(ksq));
      single &= white ? (ok_src << 8) : (ok_src >> 8);
      dbl &= white ? (ok_src << 16) : (ok_src >> 16);
    }
    c.single = single;
    c.dbl = dbl;
  }
}

CRL_HD int wg_square(int lane, int h) { return (h ? 31 : 63) - lane; }

CRL_HD void wg_lane(const Board& b, const WgCommon& c, int lane, WgLane& w) {
  u64 packed = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int sq = wg_square(lane, h);
    const u64 fb = bit(sq);
    u64 t = 0;
    if (fb & c.officers) {
      if (fb & b.bb[KNIGHT]) t = knight_attacks_sq(sq);
      else if (fb & b.bb[KING]) t = c.king_moves;
      else {
        if (fb & (b.bb[BISHOP] | b.bb[QUEEN])) t = bishop_attacks(sq, c.occ);
        if (fb & (b.bb[ROOK] | b.bb[QUEEN])) t |= rook_attacks(sq, c.occ);
      }
      t &= ~c.us;
      if (!(fb & b.bb[KING])) {
        t &= c.target;
        if (fb & c.pinned) t &= line_through(c.ksq, sq);
      }
    }
    w.t1[h] = t;
    u64 t3 = 0;
    if (fb & c.pawns) {
      t3 = pawn_attacks_set(fb, c.white) & c.them & c.target;
      if (fb & c.pinned) t3 &= line_through(c.ksq, sq);
    }
    w.t3[h] = t3;
    const u64 n1 = (u64)popc64(t);
    const u64 n3 = (u64)(popc64(t3) * ((t3 & c.promo_rank) ? 4 : 1));
    const u64 n4 = (c.single & fb) ? ((fb & c.promo_rank) ? 4u : 1u) : 0u;
    const u64 n5 = (c.dbl & fb) ? 1u : 0u;
    packed |= (n1 << (8 * h)) | (n3 << (16 + 8 * h)) | (n4 << (32 + 8 * h)) | (n5 << (48 + 8 * h));
  }
  w.packed = packed;
}

// offsets of the lane's eight groups from the inclusive prefix sum `incl` over lanes 0..lane and the warp totals `tot`;
// first = number of moves written before class 1 (king evasions when in check)
struct WgOffsets {
  int o[8];     // [A1, B1, A3, B3, A4, B4, A5, B5]
  int base6;    // where en passant starts
  int base2;    // where castling starts
};
CRL_HD void wg_offsets(u64 incl, u64 own, u64 tot, int first, int n_castle, WgOffsets& r) {
  int base = first;
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    if (f == 2) {                       // castling sits between the officer moves and the pawn captures
      r.base2 = base;
      base += n_castle;
    }
    r.o[f] = base + (int)((incl >> (8 * f)) & 255) - (int)((own >> (8 * f)) & 255);
    base += (int)((tot >> (8 * f)) & 255);
  }
  r.base6 = base;
}

CRL_HD void wg_emit(const WgCommon& c, int lane, const WgLane& w, const WgOffsets& r, u16* out) {
  const int back = c.white ? -8 : 8;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int sq = wg_square(lane, h);
    u64 t = w.t1[h];
    int o = r.o[h];
    while (t) {
      const int to = pop_msb(t);
      out[o++] = mk_move(sq, to, 0);
    }
    t = w.t3[h];
    o = r.o[2 + h];
    const bool promo = (t & c.promo_rank) != 0;
    while (t) {
      const int to = pop_msb(t);
      if (promo) {
        out[o++] = mk_move(sq, to, QUEEN);
        out[o++] = mk_move(sq, to, ROOK);
        out[o++] = mk_move(sq, to, BISHOP);
        out[o++] = mk_move(sq, to, KNIGHT);
      } else {
        out[o++] = mk_move(sq, to, 0);
      }
    }
    const u64 fb = bit(sq);
    if (c.single & fb) {
      o = r.o[4 + h];
      if (fb & c.promo_rank) {
        out[o++] = mk_move(sq + back, sq, QUEEN);
        out[o++] = mk_move(sq + back, sq, ROOK);
        out[o++] = mk_move(sq + back, sq, BISHOP);
        out[o++] = mk_move(sq + back, sq, KNIGHT);
      } else {
        out[o] = mk_move(sq + back, sq, 0);
      }
    }
    if (c.dbl & fb) out[r.o[6 + h]] = mk_move(sq + 2 * back, sq, 0);
  }
}

// the (at most ten) moves no square-owner writes: king evasions first, castling, en passant last.  Returns the number
// of en passant moves written at r.base6; *ep_legal as GenInfo.ep_legal
CRL_HD int wg_emit_rest(const Board& b, const WgCommon& c, const WgOffsets& r, u16* out, int* ep_legal) {
  if (c.in_check) {
    u64 t = c.king_moves;
    int o = 0;
    while (t) {
      const int to = pop_msb(t);
      out[o++] = mk_move(c.ksq, to, 0);
    }
  }
  {
    const int base = c.white ? 0 : 56;
    int o = r.base2;
    if (c.castle_k) out[o++] = mk_move(c.ksq, base + 6, 0);
    if (c.castle_q) out[o] = mk_move(c.ksq, base + 2, 0);
  }
  *ep_legal = 0;
  int n_ep = 0;
  const int ep = meta_ep(b.meta);
  if (!c.double_check && c.pawns && ep > 0 && !(c.occ & bit(ep))) {
    const int victim = ep + (c.white ? -8 : 8);
    const bool allowed = !c.in_check || (c.target & bit(ep)) || victim == c.checker_sq;
    if (allowed) {
      u64 cap = c.pawns & pawn_attacks_set(bit(ep), !c.white) & (c.white ? (0xFFULL << 32) : (0xFFULL << 24));
      while (cap) {
        const int from = pop_msb(cap);
        if (ep_capture_safe(b, from, ep, c.white, c.ksq)) {
          out[r.base6 + n_ep++] = mk_move(from, ep, 0);
          *ep_legal = 1;
        }
      }
    }
  }
  return n_ep;
}

#if defined(__CUDACC__)
// All 32 lanes call this with the SAME board; `out` (shared or global, >= MAX_MOVES entries) receives the legal moves in
// python-chess order.  Returns their number; *in_check / *ep_legal as GenInfo.  The caller synchronises the warp before
// reading `out`.
__device__ __forceinline__ int warp_generate_legal(const Board& b, u16* out, int lane, int* in_check, int* ep_legal) {
  WgCommon c;
  wg_common(b, c);
  *in_check = c.in_check;
  *ep_legal = 0;
  if (!c.valid) return 0;
  WgLane w;
  wg_lane(b, c, lane, w);
  u64 incl = w.packed;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const u64 v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  const u64 tot = __shfl_sync(0xffffffffu, incl, 31);
  WgOffsets r;
  wg_offsets(incl, w.packed, tot, c.in_check ? popc64(c.king_moves) : 0, c.n_castle, r);
  wg_emit(c, lane, w, r, out);
  int n_ep = 0, epl = 0;
  if (lane == 0) n_ep = wg_emit_rest(b, c, r, out, &epl);
  n_ep = __shfl_sync(0xffffffffu, n_ep, 0);
  *ep_legal = __shfl_sync(0xffffffffu, epl, 0);
  return r.base6 + n_ep;
}
#endif

}  // namespace crl
