// TEST-ONLY host build of chessrl_b200/csrc/chess_core.cuh (compiled with g++ by tests/hostsim/__init__.py).
// It exists so the rule kernels' arithmetic can be checked against the oracle in the CPU-only build
// container.  It is NOT part of libchessrl_b200.so and nothing under chessrl_b200/ loads it.
#include "../../chessrl_b200/csrc/chess_core.cuh"
#include "../../chessrl_b200/csrc/warp_gen.cuh"
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

// the warp-cooperative generator (warp_gen.cuh) with its 32 lanes run one after the other and the shuffle prefix sum
// done in a loop: same per-lane code as the device, so its move lists can be checked against the scalar generator here
int hs_movegen_warp(const uint64_t* rec, uint16_t* out, int* flags) {
  Board b;
  memcpy(b.bb, rec, 64);
  b.meta = rec[8];
  WgCommon c;
  wg_common(b, c);
  flags[0] = c.in_check;
  flags[1] = 0;
  if (!c.valid) return 0;
  WgLane w[32];
  uint64_t incl[32], run = 0;
  for (int lane = 0; lane < 32; ++lane) {
    wg_lane(b, c, lane, w[lane]);
    run += w[lane].packed;
    incl[lane] = run;
  }
  for (int i = 0; i < MAX_MOVES; ++i) out[i] = 0xEEEE;
  WgOffsets r0;
  for (int lane = 0; lane < 32; ++lane) {
    WgOffsets r;
    wg_offsets(incl[lane], w[lane].packed, run, c.in_check ? popc64(c.king_moves) : 0, c.n_castle, r);
    wg_emit(c, lane, w[lane], r, out);
    if (lane == 0) r0 = r;
  }
  int n_ep = wg_emit_rest(b, c, r0, out, &flags[1]);
  return r0.base6 + n_ep;
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

// ---- serial host run of the tree logic (tree_core.cuh) with the hash evaluator -----------------------
#include "../../chessrl_b200/csrc/tree_core.cuh"
#include "../../chessrl_b200/csrc/hash_eval.cuh"
#include <stdlib.h>
#include <utility>
#include <vector>

struct HostTree {
  Pools P;
  std::vector<int16_t> label_of;   // [5][64][64]
  std::vector<float> policy;       // one row
};

extern "C" {

enum { HS_KMAX = 64 };

void* hs_tree_new(int NN, int EA, const int16_t* label_of) {
  HostTree* t = new HostTree();
  Pools& P = t->P;
  P.G = 1; P.NN = NN; P.EA = EA; P.K = 1;
  P.g_cur = (u64*)calloc(9, 8);
  P.g_hist = (u64*)calloc(HIST_RING * 8, 8);
  P.g_keys = (u64*)calloc(KEY_RING, 8);
  P.g_moves = (u16*)calloc(MAX_GAME_PLIES, 2);
  P.g_nmoves = (int*)calloc(1, 4);
  P.g_result = (int8_t*)calloc(1, 1);
  P.g_active = (u8*)calloc(1, 1);
  P.nodes = (NodeRec*)calloc(NN, sizeof(NodeRec));
  P.g_nnodes = (int*)calloc(1, 4);
  P.g_nedges = (int*)calloc(1, 4);
  P.e_move = (u16*)calloc(EA, 2);
  P.e_prior = (float*)calloc(EA, 4);
  P.e_visits = (int*)calloc(EA, 4);
  P.e_value = (double*)calloc(EA, 8);
  P.e_child = (int*)calloc(EA, 4);
  P.e_result = (int8_t*)calloc(EA, 1);
  P.e_vloss = (u8*)calloc(EA, 1);
  P.r_visits = (int*)calloc(1, 4);
  P.r_value = (double*)calloc(1, 8);
  P.reuse = 0;
  P.nodes_prev = (NodeRec*)calloc(NN, sizeof(NodeRec));
  P.e_prior_prev = (float*)calloc(EA, 4);
  P.e_child_prev = (int*)calloc(EA, 4);
  P.g_prev_root = (int*)calloc(1, 4);
  P.g_prev_root[0] = -1;
  P.s_node = (int*)calloc(HS_KMAX, 4);
  P.s_kind = (int*)calloc(HS_KMAX, 4);
  P.s_moves = (u16*)calloc(HS_KMAX * MAX_MOVES, 2);
  P.s_nmoves = (int*)calloc(HS_KMAX, 4);
  P.s_row = (int*)calloc(HS_KMAX, 4);
  P.s_path = nullptr;      // device-only (level-parallel backup)
  P.s_depth = nullptr;
  P.s_wave_n = (int*)calloc(1, 4);
  P.g_sims_left = (int*)calloc(1, 4);
  P.eval_list = (int*)calloc(HS_KMAX, 4);
  P.eval_n = (int*)calloc(1, 4);
  P.err = (int*)calloc(1, 4);
  P.counters = (long long*)calloc(4, 8);
  t->label_of.assign(label_of, label_of + 5 * 4096);
  t->policy.resize(1968);
  return t;
}

// returns number of moves accepted
int hs_game_set(void* h, const uint64_t* start, const uint16_t* moves, int n) {
  HostTree* t = (HostTree*)h;
  Pools& P = t->P;
  for (int k = 0; k < 9; ++k) P.g_cur[k] = start[k];
  P.g_nmoves[0] = 0;
  P.g_prev_root[0] = -1;
  game_refresh(P, 0, nullptr, nullptr);
  int ok = 0;
  for (int i = 0; i < n; ++i) ok += game_move(P, 0, moves[i]);
  return ok;
}
int hs_game_move(void* h, uint16_t mv) {
  Pools& P = ((HostTree*)h)->P;
  const int ok = game_move(P, 0, mv);
  if (ok) P.g_prev_root[0] = -1;
  return ok;
}
int hs_game_result(void* h) { return ((HostTree*)h)->P.g_result[0]; }
void hs_cur_record(void* h, uint64_t* out9) {
  for (int k = 0; k < 9; ++k) out9[k] = ((HostTree*)h)->P.g_cur[k];
}

static void host_eval(HostTree* t, const u64* rec, uint64_t seed, int bits, float* value) {
  Board b = load_rec(rec);
  u64 hh = eval_hash(b, seed);
  for (int i = 0; i < 1968; ++i) t->policy[i] = hash_policy(hh, i, bits);
  *value = hash_value(hh);
}

// reuse != 0: evaluation reuse (tree_core.cuh "evaluation reuse") -- the pools swap roles first and expansions look
// their twin up in the tree of the previous hs_search call, as crl_mcts_begin_move / k_select_expand do on the device.
// Returns the number of evaluations RUN (low 24 bits) and the error flags.
static int hs_search_impl(void* h, int sims, uint64_t seed, int bits, int reuse) {
  HostTree* t = (HostTree*)h;
  Pools& P = t->P;
  float v;
  P.reuse = reuse ? 1 : 0;
  if (reuse) {
    std::swap(P.nodes, P.nodes_prev);
    std::swap(P.e_prior, P.e_prior_prev);
    std::swap(P.e_child, P.e_child_prev);
  }
  int evals = 0;
  if (root_init(P, 0, reuse != 0)) {
    host_eval(t, P.nodes[0].p2, seed, bits, &v);
    store_priors(P, 0, 0, t->policy.data(), t->label_of.data());
    ++evals;
  }
  for (int s = 0; s < sims; ++s) {
    int node, term;
    select_descend(P, 0, [&](const NodeRec& n) { return best_edge_serial(P, 0, n); }, &node, &term);
    double val;
    if (term) {
      val = (double)P.nodes[node].result;
    } else {
      int child;
      int kind = expand_child(P, 0, 0, node, &child);
      node = child;
      const int twin = P.nodes[child].prev;
      bool adopted = false;
      if (kind == KIND_NEED_REPLY) {
        int pick = -1;
        if (twin >= 0) pick = find_move(P.s_moves, P.s_nmoves[0], P.nodes_prev[twin].reply);
        if (pick < 0) {
          host_eval(t, P.nodes[child].p1, seed, bits, &v);
          ++evals;
        }
        kind = reply_child(P, 0, 0, child, t->policy.data(), t->label_of.data(), pick);
        if (pick >= 0 && kind == KIND_EVAL_LEAF) adopted = adopt_evaluation(P, 0, child, twin);
      }
      if (kind == KIND_EVAL_LEAF && adopted) {
        val = (double)P.nodes[child].v;
      } else if (kind == KIND_EVAL_LEAF) {
        host_eval(t, P.nodes[child].p2, seed, bits, &v);
        ++evals;
        store_priors(P, 0, child, t->policy.data(), t->label_of.data());
        P.nodes[child].v = v;
        P.nodes[child].evald = 1;
        val = (double)v;
      } else {
        val = (double)P.nodes[child].result;
      }
    }
    backup(P, 0, node, val);
  }
  return evals | (*P.err << 24);
}
int hs_search(void* h, int sims, uint64_t seed, int bits) { return hs_search_impl(h, sims, seed, bits, 0); }
int hs_search_reuse(void* h, int sims, uint64_t seed, int bits) { return hs_search_impl(h, sims, seed, bits, 1); }

// k_commit(apply = 1): play (our move, reply) of root child k and link the game to that child for the next search
int hs_commit(void* h, int k) {
  Pools& P = ((HostTree*)h)->P;
  const NodeRec& root = P.nodes[0];
  int next_root = -1, played = 0;
  if (k >= 0 && k < root.n_exp) {
    const int c = P.e_child[root.edge0 + k];
    const NodeRec& cn = P.nodes[c];
    u16 m0 = MOVE_NONE, m1 = MOVE_NONE;
    if (cn.reply != MOVE_NONE) {
      m0 = cn.move;
      m1 = cn.reply;
    } else if (P.g_nmoves[0] >= 1) {
      m0 = P.g_moves[P.g_nmoves[0] - 1];
      m1 = cn.move;
    }
    const int ok0 = game_move(P, 0, m0), ok1 = game_move(P, 0, m1);
    played = ok0 + ok1;
    if (ok0 && ok1 && cn.reply == m1 && cn.move == m0) next_root = c;
  }
  P.g_prev_root[0] = next_root;
  return played;
}

// wave mode (K in-flight simulations): the same three phases as k_select_wave / k_reply / k_finalize_wave, serially
int hs_search_wave(void* h, int sims, int K, uint64_t seed, int bits) {
  HostTree* t = (HostTree*)h;
  Pools& P = t->P;
  if (K > HS_KMAX) return -1;
  P.K = K;
  float v;
  root_init(P, 0);
  host_eval(t, P.nodes[0].p2, seed, bits, &v);
  store_priors(P, 0, 0, t->policy.data(), t->label_of.data());
  int evals = 1, left = sims, waves = 0;
  std::vector<double> vals(K);
  while (left > 0) {
    const int kmax = left < K ? left : K;
    int used = 0;
    for (int j = 0; j < kmax; ++j) {
      int node = 0;
      int what = select_descend_wave(P, 0, [&](const NodeRec& n) { return best_edge_serial(P, 0, n, true); }, &node);
      if (what == 2) break;
      wave_take_slot(P, 0, used, what, node);
      ++used;
    }
    for (int j = 0; j < used; ++j) {
      if (P.s_kind[j] != KIND_NEED_REPLY) continue;
      host_eval(t, P.nodes[P.s_node[j]].p1, seed, bits, &v);
      ++evals;
      P.s_kind[j] = reply_child(P, 0, j, P.s_node[j], t->policy.data(), t->label_of.data());
    }
    for (int j = 0; j < used; ++j) {
      const int node = P.s_node[j];
      double val;
      if (P.s_kind[j] == KIND_EVAL_LEAF) {
        host_eval(t, P.nodes[node].p2, seed, bits, &v);
        ++evals;
        store_priors(P, 0, node, t->policy.data(), t->label_of.data());
        val = (double)v;
      } else {
        val = (double)P.nodes[node].result;
      }
      backup(P, 0, node, val);
      vloss_add(P, 0, node, -1);
    }
    left -= used;
    ++waves;
  }
  P.K = 1;
  return waves | (*P.err << 24);
}

// root stats in child creation order
int hs_root_stats(void* h, int* visits, double* values, float* priors, uint16_t* moves, uint16_t* replies,
                  int* results, int* root_visits, double* root_value) {
  Pools& P = ((HostTree*)h)->P;
  const NodeRec& r = P.nodes[0];
  for (int k = 0; k < r.n_exp; ++k) {
    visits[k] = P.e_visits[k];
    values[k] = P.e_value[k];
    priors[k] = P.e_prior[k];
    const NodeRec& c = P.nodes[P.e_child[k]];
    moves[k] = c.move;
    replies[k] = c.reply;
    results[k] = c.result;
  }
  *root_visits = P.r_visits[0];
  *root_value = P.r_value[0];
  return r.n_exp;
}
// visits of the children of root child k
int hs_grandchild_visits(void* h, int k, int* visits) {
  Pools& P = ((HostTree*)h)->P;
  const NodeRec& c = P.nodes[P.e_child[k]];
  for (int j = 0; j < c.n_exp; ++j) visits[j] = P.e_visits[c.edge0 + j];
  return c.n_exp;
}
double hs_edge_score(int visits, double value, float prior, int child_result) {
  return edge_score(visits, value, prior, child_result);
}
// planes of (node, which) via the history walk: out[9][8] bitboards, returns how many positions exist
int hs_history(void* h, int node, int which, uint64_t* out) {
  Pools& P = ((HostTree*)h)->P;
  const NodeRec& n = P.nodes[node];
  Board b = load_rec(which == 2 ? n.p2 : n.p1);
  Cursor c{node, which, meta_ply(b.meta)};
  int cnt = 0;
  cursor_bitboards(P, 0, c, out);
  cnt = 1;
  while (cnt < 9 && cursor_prev(P, 0, c)) {
    cursor_bitboards(P, 0, c, out + 8 * cnt);
    ++cnt;
  }
  return cnt;
}
}
