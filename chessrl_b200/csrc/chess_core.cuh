// chess_core.cuh -- bitboard rules for the lockstep self-play engine (sm_100a).
//
// What this replaces (reference file:line): everything python-chess does for
// game.Game.get_legal_moves (game.py:43-57), Game.move (game.py:28-41) and Game.get_result
// (game.py:92-109).  Legal moves come out in python-chess 0.28.3 GENERATION ORDER because the tree
// expands `unexpanded_actions.pop()` (mctree.py:55-56), zips priors by that order (mctree.py:298-303) and
// breaks argmax ties by index (agentdistributed.py:58).
//
// Design: no magic / PEXT tables.  Sliding attacks are hyperbola-quintessence with the hardware bit
// reversal (BREV), line masks are computed arithmetically, pins / check masks / the enemy attack map are
// computed once per position, so the generator is pure register arithmetic: nothing to fetch but the 72-byte
// board record.  Every function is CRL_HD so the same source can be compiled by g++ into the TEST-ONLY
// host harness (tests/hostsim) that checks it against the oracle where no GPU exists; the shipped library
// only ever runs it on the device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CRL_HD __host__ __device__ __forceinline__
#else
#define CRL_HD inline
#endif

namespace crl {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

enum { PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5, OCC_W = 6, OCC_B = 7 };
enum { NO_PIECE = 7 };
enum { RESULT_NONE = 2 };           // game not over (Game.get_result() is None)
enum { MAX_MOVES = 256 };           // >= 218, the known maximum
static const u16 MOVE_NONE = 0xFFFF;

// ---- meta word layout (one u64 per board) ---------------------------------------------------------
//  bit 0       side to move (1 = white)                     chess.Board.turn
//  bits 1..4   clean castling rights K,Q,k,q                 Board.clean_castling_rights()
//  bits 5..11  ep square + 1 (0 = none); set after EVERY double push, like Board.push
//  bits 12..23 halfmove clock (saturates at 4095)
//  bits 24..37 fullmove number
//  bits 38..51 ply = len(board.move_stack)
//  bits 52..59 length of the reversible run ending here (for fivefold; saturates at 255)
struct Board {
  u64 bb[8];
  u64 meta;
};

CRL_HD int meta_turn(u64 m) { return (int)(m & 1); }
CRL_HD int meta_castle(u64 m) { return (int)((m >> 1) & 15); }
CRL_HD int meta_ep(u64 m) { return (int)((m >> 5) & 127) - 1; }   // -1 = none
CRL_HD int meta_halfmove(u64 m) { return (int)((m >> 12) & 4095); }
CRL_HD int meta_fullmove(u64 m) { return (int)((m >> 24) & 16383); }
CRL_HD int meta_ply(u64 m) { return (int)((m >> 38) & 16383); }
CRL_HD int meta_revlen(u64 m) { return (int)((m >> 52) & 255); }
CRL_HD u64 meta_pack(int turn, int castle, int ep, int half, int full, int ply, int rev) {
  if (half > 4095) half = 4095;
  if (full > 16383) full = 16383;
  if (ply > 16383) ply = 16383;
  if (rev > 255) rev = 255;
  return (u64)(turn & 1) | ((u64)(castle & 15) << 1) | ((u64)((ep + 1) & 127) << 5) | ((u64)half << 12) |
         ((u64)full << 24) | ((u64)ply << 38) | ((u64)rev << 52);
}

// ---- move word: from | to<<6 | promo<<12 (promo: 0 none, else KNIGHT=1..QUEEN=4) ------
CRL_HD u16 mk_move(int from, int to, int promo) { return (u16)(from | (to << 6) | (promo << 12)); }
CRL_HD int mv_from(u16 m) { return m & 63; }
CRL_HD int mv_to(u16 m) { return (m >> 6) & 63; }
CRL_HD int mv_promo(u16 m) { return (m >> 12) & 7; }

// ---- bit helpers ---------------------------------------------------------------------------------
CRL_HD int msb64(u64 x) {
#if defined(__CUDA_ARCH__)
  return 63 - __clzll((long long)x);
#else
  return 63 - __builtin_clzll(x);
#endif
}
CRL_HD int popc64(u64 x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
CRL_HD u64 brev64(u64 x) {
#if defined(__CUDA_ARCH__)
  return __brevll(x);
#else
  x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  return __builtin_bswap64(x);
#endif
}
CRL_HD u64 bit(int sq) { return 1ULL << sq; }
// 32-bit halves: a 64-bit count-leading-zeros / variable shift costs several 32-bit instructions on the GPU, so the
// hot loops that walk a set from its most significant bit down do it one word at a time
CRL_HD int msb32(u32 x) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}
// removes and returns the most significant set bit (x != 0)
CRL_HD int pop_msb(u64& x) {
  const u32 hi = (u32)(x >> 32);
  if (hi) {
    const int t = msb32(hi);
    x ^= (u64)(1u << t) << 32;
    return t + 32;
  }
  const int t = msb32((u32)x);
  x ^= (u64)(1u << t);
  return t;
}

static const u64 FILE_A = 0x0101010101010101ULL;
static const u64 FILE_B = FILE_A << 1;
static const u64 FILE_G = FILE_A << 6;
static const u64 FILE_H = FILE_A << 7;
static const u64 RANK_1 = 0xFFULL;
static const u64 RANK_8 = 0xFFULL << 56;
static const u64 DIAG_MAIN = 0x8040201008040201ULL;   // a1-h8
static const u64 DIAG_ANTI = 0x0102040810204080ULL;   // h1-a8
static const u64 LIGHT_SQ = 0x55AA55AA55AA55AAULL;
static const u64 DARK_SQ = 0xAA55AA55AA55AA55ULL;

CRL_HD u64 rank_mask(int sq) { return RANK_1 << (sq & 56); }
CRL_HD u64 file_mask(int sq) { return FILE_A << (sq & 7); }
CRL_HD u64 diag_mask(int sq) {
  int k = (sq >> 3) - (sq & 7);
  return k >= 0 ? (DIAG_MAIN << (8 * k)) : (DIAG_MAIN >> (-8 * k));
}
CRL_HD u64 anti_mask(int sq) {
  int k = (sq >> 3) + (sq & 7) - 7;
  return k >= 0 ? (DIAG_ANTI << (8 * k)) : (DIAG_ANTI >> (-8 * k));
}

// hyperbola quintessence along one line; `mask` is the full line through sq
CRL_HD u64 line_attacks(u64 occ, u64 mask, int sq) {
  u64 s = bit(sq);
  u64 m = mask & ~s;
  u64 o = occ & m;
  u64 fwd = o - s;
  u64 rev = brev64(o) - bit(63 - sq);
  return (fwd ^ brev64(rev)) & m;
}
CRL_HD u64 rook_attacks(int sq, u64 occ) {
  return line_attacks(occ, rank_mask(sq), sq) | line_attacks(occ, file_mask(sq), sq);
}
CRL_HD u64 bishop_attacks(int sq, u64 occ) {
  return line_attacks(occ, diag_mask(sq), sq) | line_attacks(occ, anti_mask(sq), sq);
}
// set-wise step attacks
CRL_HD u64 knight_attacks_set(u64 b) {
  u64 l1 = (b >> 1) & ~FILE_H, l2 = (b >> 2) & ~(FILE_G | FILE_H);
  u64 r1 = (b << 1) & ~FILE_A, r2 = (b << 2) & ~(FILE_A | FILE_B);
  u64 h1 = l1 | r1, h2 = l2 | r2;
  return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
CRL_HD u64 king_attacks_set(u64 b) {
  u64 a = ((b << 1) & ~FILE_A) | ((b >> 1) & ~FILE_H);
  u64 c = b | a;
  return a | (c << 8) | (c >> 8);
}
// step attacks of ONE square: the pattern around c3 / b2 slid to the square, wrap-around files cut off
// (a third of the instructions of the set-wise forms above, which have to shift in every direction)
CRL_HD u64 knight_attacks_sq(int sq) {
  constexpr u64 span = 0x0000000A1100110AULL;                            // knight on c3 (18)
  const u64 a = sq >= 18 ? span << (sq - 18) : span >> (18 - sq);
  return a & ((sq & 7) < 4 ? ~(FILE_G | FILE_H) : ~(FILE_A | FILE_B));
}
CRL_HD u64 king_attacks_sq(int sq) {
  constexpr u64 span = 0x0000000000070507ULL;                            // king on b2 (9)
  const u64 a = sq >= 9 ? span << (sq - 9) : span >> (9 - sq);
  return a & ((sq & 7) < 4 ? ~FILE_H : ~FILE_A);
}
CRL_HD u64 pawn_attacks_set(u64 b, int white) {
  return white ? (((b << 9) & ~FILE_A) | ((b << 7) & ~FILE_H)) : (((b >> 7) & ~FILE_A) | ((b >> 9) & ~FILE_H));
}

// full line through two aligned squares (0 if not aligned) and the open segment between them
CRL_HD u64 line_through(int a, int b) {
  if ((a >> 3) == (b >> 3)) return rank_mask(a);
  if ((a & 7) == (b & 7)) return file_mask(a);
  if (((a >> 3) - (a & 7)) == ((b >> 3) - (b & 7))) return diag_mask(a);
  if (((a >> 3) + (a & 7)) == ((b >> 3) + (b & 7))) return anti_mask(a);
  return 0;
}
CRL_HD u64 between(int a, int b) {
  int lo = a < b ? a : b, hi = a < b ? b : a;
  return line_through(a, b) & (bit(hi) - 1) & ~((bit(lo) << 1) - 1);
}

CRL_HD int piece_at(const Board& b, int sq) {
  u64 s = bit(sq);
  int t = NO_PIECE;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (b.bb[k] & s) t = k;
  return t;
}

// pieces of colour `by_white` attacking `sq` under occupancy `occ` (Board.attackers_mask)
CRL_HD u64 attackers_of(const Board& b, int sq, u64 occ, int by_white) {
  u64 s = bit(sq);
  u64 rq = b.bb[ROOK] | b.bb[QUEEN], bq = b.bb[BISHOP] | b.bb[QUEEN];
  u64 a = (king_attacks_set(s) & b.bb[KING]) | (knight_attacks_set(s) & b.bb[KNIGHT]) |
          (rook_attacks(sq, occ) & rq) | (bishop_attacks(sq, occ) & bq) |
          (pawn_attacks_set(s, !by_white) & b.bb[PAWN]);
  return a & b.bb[by_white ? OCC_W : OCC_B];
}

// ---- set-wise slider attacks: Kogge-Stone occluded fills --------------------------------------------
// All sliders of a set at once, one direction at a time: three doubling steps flood `gen` through the empty squares
// `pro`, one more single step lands on the first blocker.  No loop over pieces, so the lanes of a warp (which hold
// different numbers of sliders) never diverge, and within ONE direction every square is reached by at most one ray
// (from the nearest slider behind it) -- so popcounts per direction add up to the sliders' move count.
template <int DIR>   // 0 N, 1 S, 2 E, 3 W, 4 NE, 5 NW, 6 SE, 7 SW
CRL_HD u64 ray_attacks_set(u64 gen, u64 empty) {
  constexpr bool up = DIR == 0 || DIR == 2 || DIR == 4 || DIR == 5;                 // towards higher square numbers
  constexpr int sh = DIR <= 1 ? 8 : DIR <= 3 ? 1 : (DIR == 4 || DIR == 7) ? 9 : 7;
  constexpr u64 wrap = (DIR == 2 || DIR == 4 || DIR == 6) ? ~FILE_A : (DIR == 3 || DIR == 5 || DIR == 7) ? ~FILE_H : ~0ULL;
  u64 pro = empty & wrap;
  if (up) {
    gen |= pro & (gen << sh);
    pro &= pro << sh;
    gen |= pro & (gen << (2 * sh));
    pro &= pro << (2 * sh);
    gen |= pro & (gen << (4 * sh));
    return (gen << sh) & wrap;
  }
  gen |= pro & (gen >> sh);
  pro &= pro >> sh;
  gen |= pro & (gen >> (2 * sh));
  pro &= pro >> (2 * sh);
  gen |= pro & (gen >> (4 * sh));
  return (gen >> sh) & wrap;
}
// union of the attacks of rook-like / bishop-like sliders under occupancy `occ`
CRL_HD u64 rook_attacks_set(u64 rq, u64 occ) {
  const u64 e = ~occ;
  return ray_attacks_set<0>(rq, e) | ray_attacks_set<1>(rq, e) | ray_attacks_set<2>(rq, e) | ray_attacks_set<3>(rq, e);
}
CRL_HD u64 bishop_attacks_set(u64 bq, u64 occ) {
  const u64 e = ~occ;
  return ray_attacks_set<4>(bq, e) | ray_attacks_set<5>(bq, e) | ray_attacks_set<6>(bq, e) | ray_attacks_set<7>(bq, e);
}

// every square attacked by colour `by_white` under occupancy `occ`
CRL_HD u64 attack_map(const Board& b, u64 occ, int by_white) {
  u64 side = b.bb[by_white ? OCC_W : OCC_B];
  u64 a = pawn_attacks_set(b.bb[PAWN] & side, by_white) | knight_attacks_set(b.bb[KNIGHT] & side) |
          king_attacks_set(b.bb[KING] & side);
  const u64 rq = (b.bb[ROOK] | b.bb[QUEEN]) & side, bq = (b.bb[BISHOP] | b.bb[QUEEN]) & side;
  if (rq) a |= rook_attacks_set(rq, occ);
  if (bq) a |= bishop_attacks_set(bq, occ);
  return a;
}

// ---- legal move generation ------------------------------------------------------------------------
struct GenInfo {
  int in_check;
  int ep_legal;     // Board.has_legal_en_passant()
};

// Sink: void operator()(u16 move)   |  CountSink only counts (perft leaf bulk counting)
struct StoreSink {
  static constexpr bool kCounting = false;
  u16* out;
  int n;
  CRL_HD void add(int) {}
  CRL_HD void put(u16 m) { out[n++] = m; }
  CRL_HD void put_set(int from, u64 targets) {   // MSB -> LSB, one 32-bit word at a time
    u32 hi = (u32)(targets >> 32), lo = (u32)targets;
    while (hi) {
      const int t = msb32(hi);
      hi ^= 1u << t;
      out[n++] = (u16)(from | ((t + 32) << 6));
    }
    while (lo) {
      const int t = msb32(lo);
      lo ^= 1u << t;
      out[n++] = (u16)(from | (t << 6));
    }
  }
};
struct CountSink {
  static constexpr bool kCounting = true;   // only the NUMBER of moves is wanted: pawn moves are counted set-wise
  int n;
  CRL_HD void add(int k) { n += k; }
  CRL_HD void put(u16) { ++n; }
  CRL_HD void put_set(int, u64 targets) { n += popc64(targets); }
};

// true if the ep capture from `from` leaves our king safe (equals python-chess pin_mask + _ep_skewered)
CRL_HD bool ep_capture_safe(const Board& b, int from, int ep, int white, int ksq) {
  int victim = ep + (white ? -8 : 8);
  u64 occ = ((b.bb[OCC_W] | b.bb[OCC_B]) & ~bit(from) & ~bit(victim)) | bit(ep);
  u64 them = b.bb[white ? OCC_B : OCC_W] & ~bit(victim);
  u64 rq = (b.bb[ROOK] | b.bb[QUEEN]) & them, bq = (b.bb[BISHOP] | b.bb[QUEEN]) & them;
  if (rook_attacks(ksq, occ) & rq) return false;
  if (bishop_attacks(ksq, occ) & bq) return false;
  u64 k = bit(ksq);
  if (knight_attacks_sq(ksq) & b.bb[KNIGHT] & them) return false;
  if (pawn_attacks_set(k, white) & b.bb[PAWN] & them) return false;
  return true;
}

// The generator is instantiated per side to move (WHITE is a compile-time constant): lockstep lanes move the same
// colour at the same time, so the dispatch below does not diverge and every `white ? a : b` folds away.
template <class Sink, int WHITE>
CRL_HD GenInfo generate_legal_side(const Board& b, Sink& sink) {
  GenInfo info;
  info.in_check = 0;
  info.ep_legal = 0;
  constexpr int white = WHITE;
  const u64 us = b.bb[white ? OCC_W : OCC_B], them = b.bb[white ? OCC_B : OCC_W];
  const u64 occ = us | them;
  const u64 kings = b.bb[KING] & us;
  if (!kings) return info;                         // never happens in legal chess
  const int ksq = msb64(kings);
  const u64 kbit = bit(ksq);
  constexpr int base = white ? 0 : 56;

  // Board.checkers_mask() and the absolute pins (Board._slider_blockers) from ONE walk over the enemy sliders that
  // stand on a line with the king (found without occupancy): nothing in between = a checker, exactly one piece in
  // between = that piece is pinned if it is ours.  Step attackers (knight, pawn, the enemy king python-chess also
  // counts) are set-wise.  This replaces four hyperbola-quintessence line scans from the king's square.
  const u64 king_ring = king_attacks_sq(ksq);
  u64 checkers = ((knight_attacks_sq(ksq) & b.bb[KNIGHT]) | (pawn_attacks_set(kbit, white) & b.bb[PAWN]) |
                  (king_ring & b.bb[KING])) & them;
  u64 pinned = 0;
  {
    u64 rq = (b.bb[ROOK] | b.bb[QUEEN]) & them, bq = (b.bb[BISHOP] | b.bb[QUEEN]) & them;
    u64 snipers = ((rank_mask(ksq) | file_mask(ksq)) & rq) | ((diag_mask(ksq) | anti_mask(ksq)) & bq);
    while (snipers) {
      const int s = pop_msb(snipers);
      const u64 mid = between(ksq, s) & occ;
      if (!mid) checkers |= bit(s);
      else if (!(mid & (mid - 1))) pinned |= mid;
    }
    pinned &= us;
  }
  info.in_check = checkers != 0;

  // squares whose safety matters: the king's destinations and the castling paths that are otherwise clear.
  // The enemy attack map (the most expensive part of the generator) is skipped when there are none -- a king
  // boxed in by its own pieces, as in most opening positions.
  const u64 king_targets = king_ring & ~us;
  const int rights = meta_castle(b.meta) >> (white ? 0 : 2);
  bool castle_k = false, castle_q = false;
  if (!checkers && ksq == base + 4) {
    castle_k = (rights & 1) && !(occ & (0x60ULL << base));
    castle_q = (rights & 2) && !(occ & (0x0EULL << base));
  }
  u64 danger = 0;
  if (king_targets || castle_k || castle_q) danger = attack_map(b, occ ^ kbit, !white);   // king may not step here

  // evasion target mask for non-king pieces
  u64 target = ~0ULL;
  bool double_check = false;
  int checker_sq = -1;
  if (checkers) {
    if (checkers & (checkers - 1)) {
      double_check = true;
      target = 0;
    } else {
      checker_sq = msb64(checkers);
      target = between(ksq, checker_sq) | checkers;
    }
    // python-chess emits king evasions first
    sink.put_set(ksq, king_targets & ~danger);
    if (double_check) return info;
  }

  // (1) officers (and, when not in check, the king), source squares h8 -> a1
  u64 officers = us & ~b.bb[PAWN];
  if (checkers) officers &= ~kbit;
  if (Sink::kCounting) {
    // counting only: every unpinned officer set-wise.  Within one direction a square is reached by at most one ray /
    // one knight jump, so the popcounts per direction add up to the number of moves; no loop over pieces, no
    // divergence between lanes holding different material.  Pinned officers (rare) keep the per-piece path below.
    const u64 ok = ~us & target;
    if (!checkers) sink.add(popc64(king_targets & ~danger));
    const u64 free_o = officers & ~pinned & ~kbit;
    const u64 n = b.bb[KNIGHT] & free_o;
    if (n) {
      const u64 l1 = (n >> 1) & ~FILE_H, l2 = (n >> 2) & ~(FILE_G | FILE_H);
      const u64 r1 = (n << 1) & ~FILE_A, r2 = (n << 2) & ~(FILE_A | FILE_B);
      sink.add(popc64((l1 << 16) & ok) + popc64((l1 >> 16) & ok) + popc64((r1 << 16) & ok) + popc64((r1 >> 16) & ok) +
               popc64((l2 << 8) & ok) + popc64((l2 >> 8) & ok) + popc64((r2 << 8) & ok) + popc64((r2 >> 8) & ok));
    }
    const u64 empty = ~occ;
    const u64 rq = (b.bb[ROOK] | b.bb[QUEEN]) & free_o, bq = (b.bb[BISHOP] | b.bb[QUEEN]) & free_o;
    if (rq)
      sink.add(popc64(ray_attacks_set<0>(rq, empty) & ok) + popc64(ray_attacks_set<1>(rq, empty) & ok) +
               popc64(ray_attacks_set<2>(rq, empty) & ok) + popc64(ray_attacks_set<3>(rq, empty) & ok));
    if (bq)
      sink.add(popc64(ray_attacks_set<4>(bq, empty) & ok) + popc64(ray_attacks_set<5>(bq, empty) & ok) +
               popc64(ray_attacks_set<6>(bq, empty) & ok) + popc64(ray_attacks_set<7>(bq, empty) & ok));
    officers &= pinned & ~b.bb[KNIGHT];            // a pinned knight never moves; pinned sliders one by one
  }
  while (officers) {
    const int from = pop_msb(officers);
    const u64 fb = bit(from);
    u64 t;
    if (fb & b.bb[KNIGHT]) t = knight_attacks_sq(from);
    else if (fb & b.bb[KING]) t = king_targets & ~danger;
    else {
      t = 0;
      if (fb & (b.bb[BISHOP] | b.bb[QUEEN])) t = bishop_attacks(from, occ);
      if (fb & (b.bb[ROOK] | b.bb[QUEEN])) t |= rook_attacks(from, occ);
    }
    t &= ~us;
    if (!(fb & kbit)) {
      t &= target;
      if (fb & pinned) t &= line_through(ksq, from);
    }
    sink.put_set(from, t);
  }

  // (2) castling (never while in check); king side first (rook candidates scanned h -> a)
  if (castle_k && !(danger & (0x60ULL << base))) sink.put(mk_move(ksq, base + 6, 0));
  if (castle_q && !(danger & (0x0CULL << base))) sink.put(mk_move(ksq, base + 2, 0));

  const u64 pawns = b.bb[PAWN] & us;
  if (!pawns) return info;
  constexpr u64 promo_rank = white ? RANK_8 : RANK_1;

  // (3) pawn captures, source squares h8 -> a1, destinations high -> low, promotions Q R B N
  if (Sink::kCounting) {
    // counting only: unpinned pawns set-wise, pinned pawns (rare) one by one
    const u64 free_p = pawns & ~pinned;
    const u64 cl = white ? ((free_p << 7) & ~FILE_H) : ((free_p >> 9) & ~FILE_H);
    const u64 cr = white ? ((free_p << 9) & ~FILE_A) : ((free_p >> 7) & ~FILE_A);
    const u64 tl = cl & them & target, tr = cr & them & target;
    sink.add(popc64(tl) + popc64(tr) + 3 * (popc64(tl & promo_rank) + popc64(tr & promo_rank)));
    u64 src = pawns & pinned;
    while (src) {
      const int from = pop_msb(src);
      const u64 fb = bit(from);
      u64 t = pawn_attacks_set(fb, white) & them & target & line_through(ksq, from);
      sink.add(popc64(t) + 3 * popc64(t & promo_rank));
    }
  } else {
    u64 src = pawns;
    while (src) {
      const int from = pop_msb(src);
      const u64 fb = bit(from);
      u64 t = pawn_attacks_set(fb, white) & them & target;
      if (fb & pinned) t &= line_through(ksq, from);
      if (t & promo_rank) {
        while (t) {
          const int to = pop_msb(t);
          sink.put(mk_move(from, to, QUEEN));
          sink.put(mk_move(from, to, ROOK));
          sink.put(mk_move(from, to, BISHOP));
          sink.put(mk_move(from, to, KNIGHT));
        }
      } else {
        sink.put_set(from, t);
      }
    }
  }

  // (4) single pushes by destination, (5) double pushes by destination
  {
    u64 single, dbl;
    constexpr int back = white ? -8 : 8;
    if (white) {
      single = (pawns << 8) & ~occ;
      dbl = (single << 8) & ~occ & (0xFFULL << 24);
    } else {
      single = (pawns >> 8) & ~occ;
      dbl = (single >> 8) & ~occ & (0xFFULL << 32);
    }
    single &= target;
    dbl &= target;
    // a pinned pawn may only push along the king's file
    u64 pinned_pawns = pawns & pinned;
    if (pinned_pawns) {
      u64 ok_src = pawns & ~(pinned & ~file_mask(ksq));
      single &= white ? (ok_src << 8) : (ok_src >> 8);
      dbl &= white ? (ok_src << 16) : (ok_src >> 16);
    }
    if (Sink::kCounting) {
      sink.add(popc64(single) + 3 * popc64(single & promo_rank) + popc64(dbl));
    } else {
      while (single) {
        const int to = pop_msb(single);
        if ((white ? to >= 56 : to < 8)) {
          sink.put(mk_move(to + back, to, QUEEN));
          sink.put(mk_move(to + back, to, ROOK));
          sink.put(mk_move(to + back, to, BISHOP));
          sink.put(mk_move(to + back, to, KNIGHT));
        } else {
          sink.put(mk_move(to + back, to, 0));
        }
      }
      while (dbl) {
        const int to = pop_msb(dbl);
        sink.put(mk_move(to + 2 * back, to, 0));
      }
    }
  }

  // (6) en passant
  int ep = meta_ep(b.meta);
  if (ep > 0 && !(occ & bit(ep))) {
    int victim = ep + (white ? -8 : 8);
    bool allowed = !checkers || (target & bit(ep)) || victim == checker_sq;
    if (allowed) {
      u64 cap = pawns & pawn_attacks_set(bit(ep), !white) & (white ? (0xFFULL << 32) : (0xFFULL << 24));
      while (cap) {
        const int from = pop_msb(cap);
        if (ep_capture_safe(b, from, ep, white, ksq)) {
          sink.put(mk_move(from, ep, 0));
          info.ep_legal = 1;
        }
      }
    }
  }
  return info;
}

template <class Sink>
CRL_HD GenInfo generate_legal(const Board& b, Sink& sink) {
  return meta_turn(b.meta) ? generate_legal_side<Sink, 1>(b, sink) : generate_legal_side<Sink, 0>(b, sink);
}

// ---- make move (Board.push, standard chess) -------------------------------------------------------
// No dynamically indexed access to b.bb[]: the record stays in registers (a runtime index would push it to
// local memory).
CRL_HD void make_move(Board& b, u16 mv) {
  const int from = mv_from(mv), to = mv_to(mv), promo = mv_promo(mv);
  const int white = meta_turn(b.meta);
  const u64 fb = bit(from), tb = bit(to);
  int castle = meta_castle(b.meta);
  int half = meta_halfmove(b.meta) + 1, full = meta_fullmove(b.meta) + (white ? 0 : 1);
  const int ply = meta_ply(b.meta) + 1;
  const int old_ep = meta_ep(b.meta);
  int ep = -1;
  u64 us = white ? b.bb[OCC_W] : b.bb[OCC_B];
  u64 them = white ? b.bb[OCC_B] : b.bb[OCC_W];

  int pt = piece_at(b, from);
  const bool capture = (them & tb) != 0;
  const bool zeroing = pt == PAWN || capture;
  if (zeroing) half = 0;

  // Board.is_irreversible, evaluated before the move
  const int own_rights = (castle >> (white ? 0 : 2)) & 3;
  const int base = white ? 0 : 56;
  bool irreversible = zeroing || (own_rights && pt == KING) ||
                      ((own_rights & 1) && from == base + 7) || ((own_rights & 2) && from == base);

  // castling rights: any move from/to a rook home square kills that right; king move kills both
  const u64 ft = fb | tb;
  int kill = 0;
  if (ft & bit(7)) kill |= 1;
  if (ft & bit(0)) kill |= 2;
  if (ft & bit(63)) kill |= 4;
  if (ft & bit(56)) kill |= 8;
  if (pt == KING) kill |= white ? 3 : 12;
  castle &= ~kill;

  // a capture clears the target square everywhere first; then lift the mover
  if (capture) {
#pragma unroll
    for (int k = 0; k < 6; ++k) b.bb[k] &= ~tb;
    them &= ~tb;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (k == pt) b.bb[k] ^= fb;
  us ^= fb;

  if (pt == KING && (to - from == 2 || from - to == 2)) {
    // castling arrives as the king's two-square move; shift the rook as well
    int rook_from = to > from ? base + 7 : base;
    int rook_to = to > from ? base + 5 : base + 3;
    b.bb[ROOK] ^= bit(rook_from) | bit(rook_to);
    us ^= bit(rook_from) | bit(rook_to);
  } else if (pt == PAWN) {
    int diff = to - from;
    if (diff == 16 || diff == -16) {
      ep = (from + to) >> 1;
    } else if (to == old_ep && !capture && (diff == 7 || diff == 9 || diff == -7 || diff == -9)) {
      u64 vb = bit(to + (white ? -8 : 8));
      b.bb[PAWN] &= ~vb;
      them &= ~vb;
    }
  }
  if (promo) pt = promo;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (k == pt) b.bb[k] |= tb;
  us |= tb;
  b.bb[OCC_W] = white ? us : them;
  b.bb[OCC_B] = white ? them : us;

  int rev = irreversible ? 0 : meta_revlen(b.meta) + 1;
  b.meta = meta_pack(!white, castle, ep, half, full, ply, rev);
}

// ---- draw rules -----------------------------------------------------------------------------------
// Board.has_insufficient_material for both colours (is_insufficient_material)
CRL_HD bool side_insufficient(const Board& b, int white) {
  u64 ours = b.bb[white ? OCC_W : OCC_B], theirs = b.bb[white ? OCC_B : OCC_W];
  if (ours & (b.bb[PAWN] | b.bb[ROOK] | b.bb[QUEEN])) return false;
  if (ours & b.bb[KNIGHT]) return popc64(ours) <= 2 && !(theirs & ~b.bb[KING] & ~b.bb[QUEEN]);
  if (ours & b.bb[BISHOP]) {
    bool same = !(b.bb[BISHOP] & DARK_SQ) || !(b.bb[BISHOP] & LIGHT_SQ);
    return same && !b.bb[PAWN] && !b.bb[KNIGHT];
  }
  return true;
}
CRL_HD bool insufficient_material(const Board& b) { return side_insufficient(b, 1) && side_insufficient(b, 0); }

CRL_HD u64 mix64(u64 x) {   // splitmix64 finaliser (also the deterministic test evaluator's hash)
  x += 0x9E3779B97F4A7C15ULL;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// Board._transposition_key(): pieces, colours, turn, clean castling rights, ep square only if an ep capture
// is legal.  64-bit hash of the 11-tuple (collision odds over a <=100-position window ~ 1e-16).
CRL_HD u64 position_key(const Board& b, int ep_legal) {
  u64 h = 0x5D0C0FFEEULL;
#pragma unroll
  for (int k = 0; k < 8; ++k) h = mix64(h ^ b.bb[k]);
  u64 small = (b.meta & 0x1F) | (ep_legal ? (b.meta & (127ULL << 5)) : 0);
  return mix64(h ^ small);
}
// hash of the raw record as the deterministic test evaluator sees it (oracle/chessrl_oracle.py position_hash)
CRL_HD u64 eval_hash(const Board& b, u64 seed) {
  u64 h = seed;
#pragma unroll
  for (int k = 0; k < 8; ++k) h = mix64(h ^ b.bb[k]);
  return mix64(h ^ (b.meta & 0xFFF));
}

// Game.get_result (game.py:92-109) given the facts the generator already produced.
//   n_legal, in_check : from generate_legal;  reps : earlier occurrences of this position inside the
//   reversible run (the caller counts them from its history), so fivefold = reps >= 4.
// returns +1 / 0 / -1 (white's point of view) or RESULT_NONE.
CRL_HD int game_result(const Board& b, int n_legal, int in_check, int reps) {
  if (meta_halfmove(b.meta) >= 100 && n_legal > 0) return 0;             // can_claim_fifty_moves
  if (n_legal == 0) return in_check ? (meta_turn(b.meta) ? -1 : 1) : 0;   // mate / stalemate
  if (insufficient_material(b)) return 0;
  if (reps >= 4) return 0;                                               // fivefold repetition
  return RESULT_NONE;
}

}  // namespace crl
