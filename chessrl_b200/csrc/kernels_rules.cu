// kernels_rules.cu -- lockstep rule kernels: legal move generation, make-move, perft, breadth-first
// frontier expansion.  One board per thread over structure-of-arrays board batches (word k of board i at
// boards[k*n + i]) so every load of the 72-byte record is a coalesced 8-byte-per-lane access.
//
// Replaces python-chess behind game.Game.get_legal_moves / Game.move (game.py:28-57); see chess_core.cuh.
#include "engine.cuh"
#include "warp_gen.cuh"

namespace crl {

static constexpr int RULES_BLOCK = 128;

// Programmatic dependent launch (sm_90+): the plies of crl_perft_root_host form a chain of short kernels, each reading
// the control block the previous one wrote.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization a kernel is
// set up while its predecessor still runs and blocks here until that one has completed and flushed its memory --
// the launch latency between the plies overlaps instead of adding up.  A no-op for a normal launch.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// "my dependents may be set up now": once every block of this grid has said so (or exited), the next ply's blocks are
// scheduled into free slots and sit in grid_dependency_wait() until this grid is complete
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
static cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
  __device__ __forceinline__ void put_set(int from, u64 targets) {   // MSB -> LSB, one 32-bit word at a time
    u32 hi = (u32)(targets >> 32), lo = (u32)targets;
    while (hi) {
      const int t = msb32(hi);
      hi ^= 1u << t;
      put((u16)(from | ((t + 32) << 6)));
    }
    while (lo) {
      const int t = msb32(lo);
      lo ^= 1u << t;
      put((u16)(from | (t << 6)));
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
    // pad the staged row to a whole quad so the copy-out below moves only full 8-byte groups (the up to three
    // entries past the count are CRL_MOVE_NONE)
    for (int k = cnt; (k & 3) && k < MG_CAP; ++k) s_moves[warp][lane][k] = MOVE_NONE;
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
      const u32 a = *reinterpret_cast<const u32*>(src + k), b2 = *reinterpret_cast<const u32*>(src + k + 2);
      *reinterpret_cast<uint2*>(dst + k) = make_uint2(a, b2);
    }
  }
}

// test hook (crl_debug_movegen_warp): the warp-cooperative generator the tree kernels and the small perft plies use,
// one warp per board, same outputs as k_movegen
__global__ void __launch_bounds__(RULES_BLOCK) k_movegen_warp(const u64* __restrict__ boards, int n, u16* __restrict__ moves,
                                                              int* __restrict__ counts, u8* __restrict__ flags) {
  __shared__ u16 s_gen[RULES_BLOCK / 32][MAX_MOVES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * (RULES_BLOCK / 32) + warp;
  if (i >= n) return;
  const Board b = load_soa(boards, n, i);
  int chk, epl;
  const int cnt = warp_generate_legal(b, s_gen[warp], lane, &chk, &epl);
  __syncwarp();
  for (int k = lane; k < cnt; k += 32) moves[i * MAX_MOVES + k] = s_gen[warp][k];
  if (lane == 0) {
    counts[i] = cnt;
    if (flags) flags[i] = (u8)((chk ? 1 : 0) | (epl ? 2 : 0));
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

// one lane: depth-first walk below `root` with an explicit stack (depth >= 1)
__device__ __forceinline__ unsigned long long perft_lane(const Board& root, int depth, int bulk) {
  Board stack_b[PERFT_MAX_DEPTH];
  u16 stack_m[PERFT_MAX_DEPTH][MAX_MOVES];
  int stack_n[PERFT_MAX_DEPTH], stack_i[PERFT_MAX_DEPTH];
  unsigned long long total = 0;
  int level = 0;   // level L holds a position at distance L from the root; remaining plies = depth - L
  stack_b[0] = root;
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
  return total;
}

// (80 registers, 6 blocks per SM; forcing 64 registers / 8 blocks measured 3-7 % slower on B200)
__global__ void __launch_bounds__(RULES_BLOCK) k_perft(const u64* __restrict__ boards, int n, int depth, int bulk,
                                                       unsigned long long* __restrict__ nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (depth <= 0) {
    nodes[i] = 1;
    return;
  }
  nodes[i] = perft_lane(load_soa(boards, n, i), depth, bulk);
}

// ---- perft of ONE root without host round trips -----------------------------------------------------------
// Breadth-first plies (children placed with one warp-aggregated atomicAdd per warp, so their order is arbitrary --
// a perft total does not care) until the frontier holds >= min_frontier boards, then one depth-first walk per lane.
// Whether a ply still expands is decided on the device from the control block, so the host enqueues the whole
// sequence blindly: no count pass, no prefix sum, no synchronisation between plies.
//   ctl[0] boards in the current frontier   ctl[1] boards placed so far in the next one (atomic)
//   ctl[2] plies expanded                   ctl[3] overflow flag (next frontier > capacity)     ctl[4] total (atomic)
//   ctl[5] boards dealt by k_perft_pair     ctl[6] blocks of the running ply that have finished
//   ctl[9] "sharded" flag
// SHARDING over ranks (crl_perft_root_shard_host): every rank expands the same first plies; the first ply whose INPUT
// frontier holds >= shard_min boards stores only the children that belong to this rank -- shard_of(board) == shard, a hash
// of the child's record, because the atomic placement orders the frontier differently on every rank and in every call, so
// an index range would not partition it -- and from then on each rank expands and walks its own boards alone.  No board
// ever crosses a link; one all_reduce(sum) of the per-rank totals joins the counts.
// Frontiers ping-pong between buf[0] and buf[1] (SoA with stride `cap`); the current one is buf[ctl[2] & 1].
// does the next ply still expand breadth-first?  Yes while the frontier is small, or while the remaining depth is
// more than one lane's stack can walk; never beyond depth-1 plies (the last ply is always counted by the walk).
// With `pair` the last TWO plies belong to k_perft_pair (the last-but-one ply is expanded and counted in one pass,
// never stored), so breadth-first expansion stops one ply earlier.
template <class CtlPtr>
__device__ __forceinline__ bool bfs_active(CtlPtr ctl, long long min_frontier, int depth, int pair) {
  const int plies = (int)ctl[2];
  const int last = (pair && depth >= 2) ? depth - 2 : depth - 1;
  if (ctl[3] || plies >= last) return false;
  return (long long)ctl[0] < min_frontier || depth - plies > PERFT_MAX_DEPTH;
}

// Each warp takes 32 parent boards per round.  Phase 1: one lane per parent generates its legal moves into a shared-
// memory row (and parks the parent record there).  Phase 2: the warp's children are dealt round-robin to the lanes --
// child j of the warp -> lane j % 32, its parent found by a 5-step binary search over the warp's prefix sums -- so every
// lane makes one move per step and the 72-byte records of 32 consecutive children are stored with nine fully coalesced
// 256-byte writes (one lane writing all children of its own parent scatters 8-byte stores over 32 different rows).
static constexpr int BFS_CAP = 96;          // moves per parent staged in shared memory; the (rare) rest stays with its lane

struct ShardSpec {
  int shard, n_shards;
  long long shard_min;
};
static constexpr int CTL_WORDS = 16;
// which rank a board belongs to: any function of the record alone will do (transpositions land on the same rank)
__device__ __forceinline__ int shard_of(const Board& b, int n_shards) {
  const u64 x = (b.bb[OCC_W] * 0x9E3779B97F4A7C15ULL) ^ (b.bb[OCC_B] * 0xC2B2AE3D27D4EB4FULL) ^ b.bb[PAWN] ^ (b.bb[QUEEN] << 1) ^
                (b.bb[ROOK] >> 1) ^ b.meta;
  return (int)(mix64(x) % (u64)n_shards);
}

struct BfsSink {
  static constexpr bool kCounting = false;
  u16* row;        // shared-memory row of this parent
  u16* spill;      // this lane's local overflow list (moves BFS_CAP ..)
  int n;
  __device__ __forceinline__ void add(int) {}
  __device__ __forceinline__ void put(u16 m) {
    if (n < BFS_CAP) row[n] = m;
    else spill[n - BFS_CAP] = m;
    ++n;
  }
  __device__ __forceinline__ void put_set(int from, u64 targets) {   // MSB -> LSB, one 32-bit word at a time
    u32 hi = (u32)(targets >> 32), lo = (u32)targets;
    while (hi) {
      const int t = msb32(hi);
      hi ^= 1u << t;
      put((u16)(from | ((t + 32) << 6)));
    }
    while (lo) {
      const int t = msb32(lo);
      lo ^= 1u << t;
      put((u16)(from | (t << 6)));
    }
  }
};

typedef u64 (*BfsBoards)[9][32];
typedef u16 (*BfsMoves)[32][BFS_CAP];
typedef int (*BfsPre)[33];

// one breadth-first ply over parents i0, i0 + istride, ... of `in` (n boards) for the calling thread's warp
// FILTER: the ply that splits the frontier over the ranks keeps only this rank's children (compacted per round of 32
// with a ballot, one atomicAdd per round)
template <bool FILTER>
__device__ __forceinline__ void bfs_expand(const u64* __restrict__ in, u64* __restrict__ out, long long cap, long long n,
                                           unsigned long long* __restrict__ ctl, long long i0, long long istride,
                                           BfsBoards s_board, BfsMoves s_moves, BfsPre s_pre, int shard, int n_shards) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_pad = (n + 31) & ~31LL;                            // whole warps stay together
  for (long long i = i0; i < n_pad; i += istride) {
    // ---- phase 1: one parent per lane ----
    u16 spill[MAX_MOVES - BFS_CAP];
    BfsSink sink{s_moves[warp][lane], spill, 0};
    Board b;
    if (i < n) {
      b = load_soa(in, cap, i);
      generate_legal(b, sink);
#pragma unroll
      for (int k = 0; k < 8; ++k) s_board[warp][k][lane] = b.bb[k];
      s_board[warp][8][lane] = b.meta;
    }
    int pre = sink.n;                                                  // inclusive prefix sum over the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, pre, off);
      if (lane >= off) pre += v;
    }
    const int warp_total = __shfl_sync(0xffffffffu, pre, 31);
    s_pre[warp][lane] = pre - sink.n;
    if (lane == 31) s_pre[warp][32] = warp_total;
    unsigned long long base = 0;
    if (!FILTER) {
      if (lane == 0 && warp_total) base = atomicAdd(&ctl[1], (unsigned long long)warp_total);
      base = __shfl_sync(0xffffffffu, base, 0);
    }
    __syncwarp();
    if (!FILTER && base + warp_total > (unsigned long long)cap) {
      if (lane == 0) ctl[3] = 1;
      __syncwarp();
      continue;
    }
    // ---- phase 2: children dealt round-robin ----
    for (int j0 = 0; j0 < warp_total; j0 += 32) {
      const int j = j0 + lane;
      bool keep = false;
      Board c;
      if (j < warp_total) {
        int lo = 0, hi = 32;
#pragma unroll
        for (int step = 0; step < 5; ++step) {
          const int mid = (lo + hi) >> 1;
          if (s_pre[warp][mid] <= j) lo = mid;
          else hi = mid;
        }
        const int k = j - s_pre[warp][lo];
        if (k < BFS_CAP) {                                             // else: stays with its parent's lane (phase 3)
#pragma unroll
          for (int w = 0; w < 8; ++w) c.bb[w] = s_board[warp][w][lo];
          c.meta = s_board[warp][8][lo];
          make_move(c, s_moves[warp][lo][k]);
          keep = !FILTER || shard_of(c, n_shards) == shard;
        }
      }
      if (!FILTER) {
        if (keep) store_soa(out, cap, (long long)base + j, c);
      } else {
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        unsigned long long rbase = 0;
        if (lane == 0 && mask) rbase = atomicAdd(&ctl[1], (unsigned long long)__popc(mask));
        rbase = __shfl_sync(0xffffffffu, rbase, 0);
        if (rbase + __popc(mask) > (unsigned long long)cap) {
          if (lane == 0) ctl[3] = 1;
        } else if (keep) {
          store_soa(out, cap, (long long)rbase + __popc(mask & ((1u << lane) - 1)), c);
        }
      }
    }
    // ---- phase 3: parents with more than BFS_CAP moves finish their own list ----
    if (sink.n > BFS_CAP) {
      const long long o = (long long)base + pre - sink.n;
      for (int k = BFS_CAP; k < sink.n; ++k) {
        Board c = b;
        make_move(c, spill[k - BFS_CAP]);
        if (!FILTER) {
          store_soa(out, cap, o + k, c);
        } else if (shard_of(c, n_shards) == shard) {
          const unsigned long long at = atomicAdd(&ctl[1], 1ull);
          if (at + 1 > (unsigned long long)cap) ctl[3] = 1;
          else store_soa(out, cap, (long long)at, c);
        }
      }
    }
    __syncwarp();                                                      // the rows are reused by the next round
  }
}

// The same ply with ONE WARP PER PARENT (warp_gen.cuh) for small frontiers: a thread-per-parent ply costs one scalar
// move generation of latency (~30 us) however few parents there are; 32 lanes on one board bring that to a few us,
// and the children of a parent are made by consecutive lanes, so their records are stored coalesced.
//   gen = this warp's shared-memory move row (>= MAX_MOVES entries); parents w0, w0 + wstride, ...
__device__ __forceinline__ void bfs_expand_warp(const u64* __restrict__ in, u64* __restrict__ out, long long cap, long long n,
                                                unsigned long long* __restrict__ ctl, long long w0, long long wstride,
                                                u16* gen) {
  const int lane = threadIdx.x & 31;
  for (long long i = w0; i < n; i += wstride) {
    const Board b = load_soa(in, cap, i);                              // same address in every lane: one broadcast load
    int chk, epl;
    const int cnt = warp_generate_legal(b, gen, lane, &chk, &epl);
    unsigned long long base = 0;
    if (lane == 0 && cnt) base = atomicAdd(&ctl[1], (unsigned long long)cnt);
    base = __shfl_sync(0xffffffffu, base, 0);                          // (also orders the row's writes before its reads)
    __syncwarp();
    if (base + cnt > (unsigned long long)cap) {
      if (lane == 0) ctl[3] = 1;
    } else {
      for (int j = lane; j < cnt; j += 32) {
        Board c = b;
        make_move(c, gen[j]);
        store_soa(out, cap, (long long)base + j, c);
      }
    }
    __syncwarp();                                                      // the row is reused by the next parent
  }
}
static constexpr long long BFS_WARP_MAX = 12288;   // frontiers up to this size expand one warp per parent

// does the ply that expands the current frontier split it over the ranks?  (uniform: read before anybody commits)
template <class CtlPtr>
__device__ __forceinline__ bool bfs_splits(CtlPtr ctl, const ShardSpec& sh) {
  return sh.n_shards > 1 && !ctl[9] && (long long)ctl[0] >= sh.shard_min;
}
// the frontier just written becomes the current one
__device__ __forceinline__ void bfs_commit(unsigned long long* ctl, const ShardSpec& sh) {
  const bool split = bfs_splits(ctl, sh);
  ctl[0] = ctl[1];
  ctl[1] = 0;
  ctl[2] += 1;
  if (split) ctl[9] = 1;
}

// One grid-wide ply.  The LAST block to finish commits the ply (ctl[6] counts finished blocks), so a ply is one launch.
__global__ void __launch_bounds__(RULES_BLOCK) k_bfs_ply(u64* __restrict__ buf0, u64* __restrict__ buf1, long long cap,
                                                         unsigned long long* __restrict__ ctl, long long min_frontier,
                                                         int depth, int pair, ShardSpec sh) {
  __shared__ u64 s_board[RULES_BLOCK / 32][9][32];
  __shared__ __align__(8) u16 s_moves[RULES_BLOCK / 32][32][BFS_CAP];
  __shared__ int s_pre[RULES_BLOCK / 32][33];
  grid_launch_dependents();
  grid_dependency_wait();                                              // launched early (programmatic dependent launch)
  if (!bfs_active(ctl, min_frontier, depth, pair)) return;            // uniform for the whole grid: nobody commits before
  const long long n = (long long)ctl[0];                              // every block has read the control block
  const int plies = (int)ctl[2];
  const u64* in = (plies & 1) ? buf1 : buf0;
  u64* out = (plies & 1) ? buf0 : buf1;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, tstride = (long long)gridDim.x * blockDim.x;
  if (bfs_splits(ctl, sh))
    bfs_expand<true>(in, out, cap, n, ctl, t0, tstride, s_board, s_moves, s_pre, sh.shard, sh.n_shards);
  else if (n <= BFS_WARP_MAX)
    bfs_expand_warp(in, out, cap, n, ctl, (long long)blockIdx.x * (RULES_BLOCK / 32) + (threadIdx.x >> 5),
                    (long long)gridDim.x * (RULES_BLOCK / 32), &s_moves[threadIdx.x >> 5][0][0]);
  else
    bfs_expand<false>(in, out, cap, n, ctl, t0, tstride, s_board, s_moves, s_pre, 0, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl[6], 1ull) + 1 == (unsigned long long)gridDim.x) {
      ctl[6] = 0;
      bfs_commit(ctl, sh);
    }
  }
}

// The root and the first plies in ONE block (1 -> 20 -> 400 boards from the start position): the root record arrives as
// a kernel parameter, the control block is initialised here, and up to `n_plies` plies run back to back with a block
// barrier in between -- no copies, no memset, no launch per tiny ply.
struct RootRecord {
  u64 w[9];
};
static constexpr int FIRST_THREADS = 512;          // 16 warps: the 20 children of the start position in two rounds
static constexpr long long FIRST_MAX = 256;        // k_bfs_first runs its (at most two) plies whatever the root's move count
                                                   // (<= 218), so the host knows which plies are left for the grid kernels
__global__ void __launch_bounds__(FIRST_THREADS) k_bfs_first(RootRecord root, u64* __restrict__ buf0, u64* __restrict__ buf1,
                                                             long long cap, unsigned long long* __restrict__ ctl,
                                                             long long min_frontier, int depth, int pair, int n_plies,
                                                             ShardSpec sh) {
  __shared__ u16 s_gen[FIRST_THREADS / 32][MAX_MOVES];
  grid_launch_dependents();
  if (threadIdx.x < CTL_WORDS) ctl[threadIdx.x] = threadIdx.x == 0 ? 1ull : 0ull;
  if (threadIdx.x < 9) buf0[(long long)threadIdx.x * cap] = root.w[threadIdx.x];     // column 0 of the SoA buffer
  __syncthreads();
  const volatile unsigned long long* vctl = ctl;
  for (int ply = 0; ply < n_plies; ++ply) {
    if (!bfs_active(vctl, min_frontier, depth, pair)) break;          // uniform: same control block for every thread
    const long long n = (long long)vctl[0];
    if (n > FIRST_MAX) break;                                          // the grid-wide ply kernels take over
    const int plies = (int)vctl[2];
    const u64* in = (plies & 1) ? buf1 : buf0;
    u64* out = (plies & 1) ? buf0 : buf1;
    __syncthreads();
    bfs_expand_warp(in, out, cap, n, ctl, threadIdx.x >> 5, FIRST_THREADS / 32, s_gen[threadIdx.x >> 5]);
    __syncthreads();
    if (threadIdx.x == 0) bfs_commit(ctl, sh);
    __syncthreads();
  }
}
__global__ void __launch_bounds__(RULES_BLOCK) k_perft_walk(const u64* __restrict__ buf0, const u64* __restrict__ buf1,
                                                            long long cap, unsigned long long* __restrict__ ctl, int depth,
                                                            int bulk, int pair, ShardSpec sh) {
  grid_dependency_wait();
  if (ctl[3]) return;
  const long long n = (long long)ctl[0];
  const bool split_here = sh.n_shards > 1 && !ctl[9];     // the frontier never grew to shard_min boards: split it now
  const int plies = (int)ctl[2];
  const u64* in = (plies & 1) ? buf1 : buf0;
  const int remaining = depth - plies;
  if (pair && remaining == 2) return;                                  // k_perft_pair counts these
  unsigned long long mine = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const Board b = load_soa(in, cap, i);
    if (split_here && shard_of(b, sh.n_shards) != sh.shard) continue;
    mine += remaining <= 0 ? 1ull : perft_lane(b, remaining, bulk);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&ctl[4], mine);
}

// OPTIONAL (CRL_PERFT_PAIR=5 / 6; off by default -- measured, it does not pay: profiles/r01_perft_pair_probe.json).
// The last two plies in one pass, nothing stored: when the frontier sits two plies above the leaves, a warp takes P of
// its boards per round (P = 32 for a large frontier, fewer for a small one so that every SM still gets warps), lanes
// 0..P-1 generate their board's moves into shared-memory rows exactly as k_bfs_ply does, and the warp's children are
// dealt round-robin to all 32 lanes, which make the move in registers and COUNT the child's legal moves (bulk) or
// generate and make each of them (every leaf made).  Against "expand the last-but-one ply, then walk it" this drops
// the 72-byte store and the 72-byte load of every board of the largest frontier (Kiwipete depth 5: 4.1 M boards =
// 588 MB; start depth 7: 119 M boards = 17 GB) and one ply + commit launch pair, while every lane still does one
// make + one generation per step, so warps stay converged.   ctl[5] = boards dealt (the lockstep lanes of the walk).
// Result on B200: bit-identical totals, but 0-8 % SLOWER with leaf bulk counting and 5-15 % slower making every leaf
// (Kiwipete d5 0.68 -> 0.71 ms, start d7 11.6 -> 11.7 ms at 96 registers; worse at 80): the stored ply's HBM traffic
// was already hidden behind the integer pipe, and the fused kernel runs at 20 warps per SM where the walk runs at 28.
template <int BULK>
__device__ __forceinline__ unsigned long long pair_leaf(const Board& c) {
  if (BULK) {
    CountSink cs{0};
    generate_legal(c, cs);
    return (unsigned long long)cs.n;
  }
  u16 local[MAX_MOVES];
  StoreSink ss{local, 0};
  generate_legal(c, ss);
  unsigned long long t = 0;
  for (int k = 0; k < ss.n; ++k) {
    Board d = c;
    make_move(d, local[k]);
    t += (d.bb[KING] != 0);
  }
  return t;
}

// MINB = resident blocks per SM the register allocation aims at: 5 -> 96 registers, no spills, 20 warps per SM;
// 6 -> 80 registers, ~130 bytes of spills, 24 warps per SM (CRL_PERFT_PAIR=5 / 6 picks; see scripts/perft_pair_probe.py)
template <int BULK, int MINB>
__global__ void __launch_bounds__(RULES_BLOCK, MINB) k_perft_pair(const u64* __restrict__ buf0, const u64* __restrict__ buf1,
                                                            long long cap, unsigned long long* __restrict__ ctl, int depth) {
  __shared__ u64 s_board[RULES_BLOCK / 32][9][32];
  __shared__ __align__(8) u16 s_moves[RULES_BLOCK / 32][32][BFS_CAP];
  __shared__ int s_pre[RULES_BLOCK / 32][33];
  if (ctl[3]) return;
  const long long n = (long long)ctl[0];
  const int plies = (int)ctl[2];
  if (depth - plies != 2) return;                                      // uniform for the whole grid
  const u64* in = (plies & 1) ? buf1 : buf0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_warps = (long long)gridDim.x * (RULES_BLOCK / 32);
  // boards per warp and round: 32 when that still leaves >= 32 warp-rounds per SM, else fewer (phase 1 then runs on
  // P of the 32 lanes, but every SM gets work: Kiwipete depth 5 deals 97,862 boards as 6,116 rounds of 16)
  const long long target = n_warps < 148 * 32 ? n_warps : 148 * 32;
  int P = 32;
  while (P > 1 && n < (long long)P * target) P >>= 1;
  unsigned long long mine = 0, dealt = 0;
  for (long long r = (long long)blockIdx.x * (RULES_BLOCK / 32) + warp; r * P < n; r += n_warps) {
    // ---- phase 1: lanes 0..P-1 take one board each ----
    const long long i = r * P + lane;
    u16 spill[MAX_MOVES - BFS_CAP];
    BfsSink sink{s_moves[warp][lane], spill, 0};
    if (lane < P && i < n) {
      const Board b = load_soa(in, cap, i);
      generate_legal(b, sink);
#pragma unroll
      for (int k = 0; k < 8; ++k) s_board[warp][k][lane] = b.bb[k];
      s_board[warp][8][lane] = b.meta;
    }
    int pre = sink.n;                                                  // inclusive prefix sum over the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, pre, off);
      if (lane >= off) pre += v;
    }
    const int warp_total = __shfl_sync(0xffffffffu, pre, 31);
    s_pre[warp][lane] = pre - sink.n;
    if (lane == 31) s_pre[warp][32] = warp_total;
    __syncwarp();
    if (lane == 0) dealt += (unsigned long long)warp_total;
    // ---- phase 2: children dealt round-robin, made and counted in registers ----
    for (int j = lane; j < warp_total; j += 32) {
      int lo = 0, hi = 32;
#pragma unroll
      for (int step = 0; step < 5; ++step) {
        const int mid = (lo + hi) >> 1;
        if (s_pre[warp][mid] <= j) lo = mid;
        else hi = mid;
      }
      const int k = j - s_pre[warp][lo];
      if (k >= BFS_CAP) continue;                                      // stays with its board's lane (phase 3)
      Board c;
#pragma unroll
      for (int w = 0; w < 8; ++w) c.bb[w] = s_board[warp][w][lo];
      c.meta = s_board[warp][8][lo];
      make_move(c, s_moves[warp][lo][k]);
      mine += pair_leaf<BULK>(c);
    }
    // ---- phase 3: boards with more than BFS_CAP moves finish their own list ----
    if (sink.n > BFS_CAP) {                                            // (the board is re-read from its parked copy:
      for (int k = BFS_CAP; k < sink.n; ++k) {                         //  keeping it in registers across phase 2 costs occupancy)
        Board c;
#pragma unroll
        for (int w = 0; w < 8; ++w) c.bb[w] = s_board[warp][w][lane];
        c.meta = s_board[warp][8][lane];
        make_move(c, spill[k - BFS_CAP]);
        mine += pair_leaf<BULK>(c);
      }
    }
    __syncwarp();                                                      // the rows are reused by the next round
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if (lane == 0) {
    if (mine) atomicAdd(&ctl[4], mine);
    if (dealt) atomicAdd(&ctl[5], dealt);
  }
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
int launch_movegen_warp(crl_engine_impl* e, const u64* boards, int n, u16* moves, int* counts, u8* flags) {
  if (n <= 0) return CRL_OK;
  LaunchScope ls(e, KC_MOVEGEN);
  k_movegen_warp<<<div_up(n, RULES_BLOCK / 32), RULES_BLOCK, 0, e->stream>>>(boards, n, moves, counts, flags);
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
// enqueues the whole device-driven perft of `root` (AoS record, host): the first kernel stores it and initialises ctl
int launch_perft_root(crl_engine_impl* e, const u64* root, u64* buf0, u64* buf1, long long cap, unsigned long long* ctl,
                      int depth, int bulk, long long min_frontier, int pair, int shard, int n_shards, long long shard_min) {
  const ShardSpec sh{shard, n_shards, shard_min};
  if (n_shards > 1) pair = 0;            // the optional two-ply pass is a single-GPU experiment
  if (depth < 0 || depth > 64) {
    set_error("crl_perft_root_host: depth %d is not supported", depth);
    return CRL_EINVAL;
  }
  if (depth < 2) pair = 0;
  const int max_plies = depth - 1 - (pair ? 1 : 0) > 0 ? depth - 1 - (pair ? 1 : 0) : 0;
  const int max_grid = 148 * 12;
  // the first kernel expands the root and its children (plies 1 and 2: at most 218 boards go in); the grid kernels take
  // the plies after that -- no launch that turns out to be a no-op (it cost 3 us + a launch gap per call)
  const int first_plies = max_plies < 2 ? max_plies : 2;
  RootRecord rr;
  for (int k = 0; k < 9; ++k) rr.w[k] = root[k];
  {
    LaunchScope ls(e, KC_MOVEGEN);
    k_bfs_first<<<1, FIRST_THREADS, 0, e->stream>>>(rr, buf0, buf1, cap, ctl, min_frontier, depth, pair, first_plies, sh);
    CRL_CUDA(cudaGetLastError());
  }
  // How many boards a ply holds is only known on the device, so every grid ply gets the full grid: blocks without
  // work leave after reading the control block.
  long long bound = 1;
  for (int ply = 0; ply < first_plies; ++ply) bound = bound * 218 < cap ? bound * 218 : cap;
  for (int ply = first_plies; ply < max_plies; ++ply) {
    {
      LaunchScope ls(e, KC_MOVEGEN);
      if (e->perft_pdl)
        CRL_CUDA(launch_dependent(k_bfs_ply, dim3(max_grid), dim3(RULES_BLOCK), e->stream, buf0, buf1, cap, ctl, min_frontier,
                                  depth, pair, sh));
      else
        k_bfs_ply<<<max_grid, RULES_BLOCK, 0, e->stream>>>(buf0, buf1, cap, ctl, min_frontier, depth, pair, sh);
      CRL_CUDA(cudaGetLastError());
    }
    bound = bound * 218 < cap ? bound * 218 : cap;
  }
  LaunchScope ls(e, KC_MOVEGEN, pair ? 2 : 1);
  const int grid = (int)(div_up(bound, RULES_BLOCK) < max_grid * 4 ? div_up(bound, RULES_BLOCK) : max_grid * 4);
  if (pair) {
    // one warp per board while the frontier is small (the kernel picks the boards per warp from the real count)
    const long long want = div_up(bound, (long long)(RULES_BLOCK / 32));
    const int pgrid = (int)(want < 148 * 16 ? want : 148 * 16);        // 9,472 warps: two per resident warp slot
    if (pair == 5) {
      if (bulk) k_perft_pair<1, 5><<<pgrid, RULES_BLOCK, 0, e->stream>>>(buf0, buf1, cap, ctl, depth);
      else k_perft_pair<0, 5><<<pgrid, RULES_BLOCK, 0, e->stream>>>(buf0, buf1, cap, ctl, depth);
    } else {
      if (bulk) k_perft_pair<1, 6><<<pgrid, RULES_BLOCK, 0, e->stream>>>(buf0, buf1, cap, ctl, depth);
      else k_perft_pair<0, 6><<<pgrid, RULES_BLOCK, 0, e->stream>>>(buf0, buf1, cap, ctl, depth);
    }
  }
  if (e->perft_pdl && !pair)
    CRL_CUDA(launch_dependent(k_perft_walk, dim3(grid), dim3(RULES_BLOCK), e->stream, (const u64*)buf0, (const u64*)buf1, cap,
                              ctl, depth, bulk, pair, sh));
  else
    k_perft_walk<<<grid, RULES_BLOCK, 0, e->stream>>>(buf0, buf1, cap, ctl, depth, bulk, pair, sh);
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
