// kernels_rules.cu -- lockstep rule kernels: legal move generation, make-move, perft, breadth-first
// frontier expansion.  One board per thread over structure-of-arrays board batches (word k of board i at
// boards[k*n + i]) so every load of the 72-byte record is a coalesced 8-byte-per-lane access.
//
// Replaces python-chess behind game.Game.get_legal_moves / Game.move (game.py:28-57); see chess_core.cuh.
#include "engine.cuh"

namespace crl {

static constexpr int RULES_BLOCK = 128;

// ---- movegen: boards -> move lists -------------------------------------------------------------------
// Moves are generated straight into a shared-memory row per board (rows of 33 words: lanes writing their k-th move hit
// 32 different banks) and copied out by QUADS of lanes -- 4 lanes x 8 bytes = one 32-byte sector of a board's row per
// step, 8 boards per pass -- so the [n][256] output rows are written in whole sectors instead of one scattered 2-byte
// store per move.  Moves beyond the 64 staged ones (rare) go directly to the global row.
static constexpr int MG_CAP = 64;
static constexpr int MG_ROW = MG_CAP + 2;

struct StageSink {
  static constexpr bool kCounting = false;
  u16* row;    // shared-memory row of this board
  u16* grow;   // its global row
  int n;
  __device__ __forceinline__ void add(int) {}
  __device__ __forceinline__ void put(u16 m) {
    if (n < MG_CAP) row[n] = m;
    else grow[n] = m;
    ++n;
  }
  __device__ __forceinline__ void put_set(int from, u64 targets) {   // MSB -> LSB
    while (targets) {
      int t = msb64(targets);
      targets ^= bit(t);
      put(mk_move(from, t, 0));
    }
  }
};

__global__ void __launch_bounds__(RULES_BLOCK) k_movegen(const u64* __restrict__ boards, int n,
                                                         u16* __restrict__ moves, int* __restrict__ counts,
                                                         u8* __restrict__ flags) {
  __shared__ __align__(8) u16 s_moves[RULES_BLOCK / 32][32][MG_ROW];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt = 0;
  if (i < n) {
    Board b = load_soa(boards, n, i);
    StageSink sink{s_moves[warp][lane], moves + (long long)i * MAX_MOVES, 0};
    GenInfo gi = generate_legal(b, sink);
    cnt = sink.n;
    counts[i] = cnt;
    if (flags) flags[i] = (u8)((gi.in_check ? 1 : 0) | (gi.ep_legal ? 2 : 0));
  }
  __syncwarp();
  const int base = blockIdx.x * blockDim.x + warp * 32;
  const int sub = lane & 3, bsel = lane >> 2;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int l = pass * 8 + bsel;                       // the board (lane) this quad copies
    int c = __shfl_sync(0xffffffffu, cnt, l);
    c = c < MG_CAP ? c : MG_CAP;
    if (base + l >= n) continue;
    const u16* src = s_moves[warp][l];
    u16* dst = moves + (long long)(base + l) * MAX_MOVES;
    for (int k = sub * 4; k < c; k += 16) {
      if (k + 4 <= c) {
        const u32 a = *reinterpret_cast<const u32*>(src + k), b2 = *reinterpret_cast<const u32*>(src + k + 2);
        *reinterpret_cast<uint2*>(dst + k) = make_uint2(a, b2);
      } else {
        for (int j = k; j < c; ++j) dst[j] = src[j];
      }
    }
  }
}

__global__ void __launch_bounds__(RULES_BLOCK) k_make(u64* __restrict__ boards, int n,
                                                      const u16* __restrict__ moves) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u16 mv = moves[i];
  if (mv == MOVE_NONE) return;
  Board b = load_soa(boards, n, i);
  make_move(b, mv);
  store_soa(boards, n, i, b);
}

// ---- perft: depth-first per lane with an explicit stack ----------------------------------------------
static constexpr int PERFT_MAX_DEPTH = 8;

// (80 registers, 6 blocks per SM; forcing 64 registers / 8 blocks measured 3-7 % slower on B200)
__global__ void __launch_bounds__(RULES_BLOCK) k_perft(const u64* __restrict__ boards, int n, int depth, int bulk,
                                                       unsigned long long* __restrict__ nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (depth <= 0) {
    nodes[i] = 1;
    return;
  }
  Board stack_b[PERFT_MAX_DEPTH];
  u16 stack_m[PERFT_MAX_DEPTH][MAX_MOVES];
  int stack_n[PERFT_MAX_DEPTH], stack_i[PERFT_MAX_DEPTH];
  unsigned long long total = 0;
  int level = 0;   // level L holds a position at distance L from the root; remaining plies = depth - L
  stack_b[0] = load_soa(boards, n, i);
  stack_n[0] = -1;
  while (level >= 0) {
    const int remaining = depth - level;
    if (stack_n[level] < 0) {   // first visit of this position
      if (remaining == 1) {
        if (bulk) {
          CountSink c{0};
          generate_legal(stack_b[level], c);
          total += (unsigned long long)c.n;
        } else {
          // no bulk counting: make every last-ply move as a plain perft would
          StoreSink s{stack_m[level], 0};
          generate_legal(stack_b[level], s);
          for (int k = 0; k < s.n; ++k) {
            Board c = stack_b[level];
            make_move(c, stack_m[level][k]);
            total += (c.bb[KING] != 0);
          }
        }
        --level;
        continue;
      }
      StoreSink s{stack_m[level], 0};
      generate_legal(stack_b[level], s);
      stack_n[level] = s.n;
      stack_i[level] = 0;
    }
    if (stack_i[level] >= stack_n[level]) {
      --level;
      continue;
    }
    u16 mv = stack_m[level][stack_i[level]++];
    stack_b[level + 1] = stack_b[level];
    make_move(stack_b[level + 1], mv);
    stack_n[level + 1] = -1;
    ++level;
  }
  nodes[i] = total;
}

// ---- one breadth-first ply ----------------------------------------------------------------------------
__global__ void __launch_bounds__(RULES_BLOCK) k_frontier(const u64* __restrict__ boards, int n,
                                                          const long long* __restrict__ offsets,
                                                          u64* __restrict__ out, long long out_n,
                                                          int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Board b = load_soa(boards, n, i);
  if (out == nullptr) {
    CountSink c{0};
    generate_legal(b, c);
    counts[i] = c.n;
    return;
  }
  u16 local[MAX_MOVES];
  StoreSink s{local, 0};
  generate_legal(b, s);
  if (counts) counts[i] = s.n;
  long long o = offsets[i];
  for (int k = 0; k < s.n; ++k) {
    if (o + k >= out_n) break;
    Board c = b;
    make_move(c, local[k]);
    store_soa(out, out_n, o + k, c);
  }
}

int launch_movegen(crl_engine_impl* e, const u64* boards, int n, u16* moves, int* counts, u8* flags) {
  if (n <= 0) return CRL_OK;
  LaunchScope ls(e, KC_MOVEGEN);
  k_movegen<<<div_up(n, RULES_BLOCK), RULES_BLOCK, 0, e->stream>>>(boards, n, moves, counts, flags);
  CRL_CUDA(cudaGetLastError());
  return CRL_OK;
}
int launch_make(crl_engine_impl* e, u64* boards, int n, const u16* moves) {
  if (n <= 0) return CRL_OK;
  LaunchScope ls(e, KC_MOVEGEN);
  k_make<<<div_up(n, RULES_BLOCK), RULES_BLOCK, 0, e->stream>>>(boards, n, moves);
  CRL_CUDA(cudaGetLastError());
  return CRL_OK;
}
int launch_perft(crl_engine_impl* e, const u64* boards, int n, int depth, int bulk, unsigned long long* nodes) {
  if (n <= 0) return CRL_OK;
  if (depth > PERFT_MAX_DEPTH) {
    set_error("crl_perft: depth %d exceeds the per-lane stack (%d)", depth, PERFT_MAX_DEPTH);
    return CRL_EINVAL;
  }
  LaunchScope ls(e, KC_MOVEGEN);
  k_perft<<<div_up(n, RULES_BLOCK), RULES_BLOCK, 0, e->stream>>>(boards, n, depth, bulk, nodes);
  CRL_CUDA(cudaGetLastError());
  return CRL_OK;
}
int launch_frontier(crl_engine_impl* e, const u64* boards, int n, const long long* offsets, u64* out,
                    long long out_n, int* counts) {
  if (n <= 0) return CRL_OK;
  LaunchScope ls(e, KC_MOVEGEN);
  k_frontier<<<div_up(n, RULES_BLOCK), RULES_BLOCK, 0, e->stream>>>(boards, n, offsets, out, out_n, counts);
  CRL_CUDA(cudaGetLastError());
  return CRL_OK;
}

}  // namespace crl
