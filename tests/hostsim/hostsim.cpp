// TEST-ONLY host build of chessrl_b200/csrc/chess_core.cuh (compiled with g++ by tests/hostsim/__init__.py).
// It exists so the rule kernels' arithmetic can be checked against the oracle in the CPU-only build
// container.  It is NOT part of libchessrl_b200.so and nothing under chessrl_b200/ loads it.
#include "../../chessrl_b200/csrc/chess_core.cuh"
#include <string.h>
using namespace crl;

extern "C" {

int hs_movegen(const uint64_t* rec, uint16_t* out, int* flags) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  StoreSink s{out, 0};
  GenInfo gi = generate_legal(b, s);
  CountSink c{0};
  generate_legal(b, c);
  if (c.n != s.n) return -1;
  flags[0] = gi.in_check;
  flags[1] = gi.ep_legal;
  return s.n;
}

void hs_make(uint64_t* rec, uint16_t mv) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  make_move(b, mv);
  memcpy(rec, b.bb, 64);
  rec[8] = b.meta;
}

uint64_t hs_key(const uint64_t* rec, int ep_legal) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  return position_key(b, ep_legal);
}

uint64_t hs_eval_hash(const uint64_t* rec, uint64_t seed) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  return eval_hash(b, seed);
}

int hs_result(const uint64_t* rec, int n_legal, int in_check, int reps) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  return game_result(b, n_legal, in_check, reps);
}

static uint64_t perft_rec(const Board& b, int depth, int bulk) {
  if (depth == 1 && bulk) {
    CountSink c{0};
    generate_legal(b, c);
    return (uint64_t)c.n;
  }
  uint16_t mv[MAX_MOVES];
  StoreSink s{mv, 0};
  generate_legal(b, s);
  if (depth == 1) return (uint64_t)s.n;
  uint64_t n = 0;
  for (int i = 0; i < s.n; ++i) {
    Board c = b;
    make_move(c, mv[i]);
    n += perft_rec(c, depth - 1, bulk);
  }
  return n;
}

uint64_t hs_perft(const uint64_t* rec, int depth, int bulk) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  return depth <= 0 ? 1 : perft_rec(b, depth, bulk);
}
}
