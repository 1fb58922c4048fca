"""Engine: thin Python host object over one crl_engine handle (one per process and GPU).

PyTorch is used for device memory (tensors handed to the C ABI as raw pointers) and the current CUDA
stream; every computation happens inside libchessrl_b200.so.  If CUDA or the library is missing this module
raises -- there is no CPU path.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from . import boards as B
from ._lib import check, vp


def _ptr(t):
    return vp(t.data_ptr()) if t is not None else None


def _np(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


class Engine:
    """max_games lockstep lanes, max_nodes tree nodes per game (>= simulations per move + 1), max_inflight
    in-flight simulations per game (the reference's `threads`; 1 = the exact threads=1 schedule only)."""

    def __init__(self, max_games=1, max_nodes=1024, avg_moves=64, device=None, max_inflight=1):
        if not torch.cuda.is_available():
            raise RuntimeError("chessrl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.max_games = int(max_games)
        self.max_nodes = int(max_nodes)
        self.max_inflight = int(max_inflight)
        self.reuse = False               # evaluation reuse across consecutive searches (set_reuse)
        h = vp()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            check(self.lib.crl_create_ex(ctypes.byref(h), self.device_index, self.max_games, self.max_nodes,
                                         int(avg_moves), self.max_inflight, vp(stream)))
        self.h = h
        self.weights_loaded = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.crl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----------------------------------------------------------------------------------------
    def boards_to_device(self, records):
        """records: uint64 [n, 9] (AoS, host) -> device tensor int64 [9, n] (SoA)."""
        rec = np.ascontiguousarray(np.asarray(records, dtype=np.uint64).reshape(-1, 9))
        t = torch.from_numpy(rec.view(np.int64).T.copy())
        return t.to(self.device)

    @staticmethod
    def boards_to_host(boards_t):
        return np.ascontiguousarray(boards_t.cpu().numpy().T).view(np.uint64)

    # ---- rules --------------------------------------------------------------------------------------------
    def movegen(self, boards_t):
        n = boards_t.shape[1]
        moves = torch.empty((n, B.MAX_MOVES), dtype=torch.int16, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        flags = torch.empty((n,), dtype=torch.uint8, device=self.device)
        check(self.lib.crl_movegen(self.h, _ptr(boards_t), n, _ptr(moves), _ptr(counts), _ptr(flags)))
        return moves, counts, flags

    def movegen_warp(self, boards_t):
        """Test hook: movegen() through the warp-cooperative generator (crl_debug_movegen_warp)."""
        n = boards_t.shape[1]
        moves = torch.empty((n, B.MAX_MOVES), dtype=torch.int16, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        flags = torch.empty((n,), dtype=torch.uint8, device=self.device)
        check(self.lib.crl_debug_movegen_warp(self.h, _ptr(boards_t), n, _ptr(moves), _ptr(counts), _ptr(flags)))
        return moves, counts, flags

    def make_moves(self, boards_t, moves_t):
        check(self.lib.crl_make_moves(self.h, _ptr(boards_t), boards_t.shape[1], _ptr(moves_t)))
        return boards_t

    def perft(self, boards_t, depth, bulk=True):
        n = boards_t.shape[1]
        nodes = torch.empty((n,), dtype=torch.int64, device=self.device)
        check(self.lib.crl_perft(self.h, _ptr(boards_t), n, int(depth), int(bool(bulk)), _ptr(nodes)))
        return nodes

    def perft_root(self, record, depth, bulk=True, min_frontier=1 << 20, shard=0, n_shards=1, shard_min_frontier=1 << 16):
        """perft(depth) of one position in ONE call: device-side breadth-first plies to >= min_frontier boards, then a
        depth-first walk per lane.  Returns (total, lanes, bfs_plies).  With n_shards > 1 the call computes shard
        `shard`'s part only (the frontier is split once it holds >= shard_min_frontier boards); the shards' totals add up
        to perft(depth)."""
        rec = np.ascontiguousarray(np.asarray(record, dtype=np.uint64).reshape(9))
        total = ctypes.c_uint64(0)
        lanes = ctypes.c_int64(0)
        plies = ctypes.c_int32(0)
        check(self.lib.crl_perft_root_shard_host(self.h, _np(rec, ctypes.c_uint64), int(depth), int(bool(bulk)),
                                                 int(min_frontier), int(shard), int(n_shards), int(shard_min_frontier),
                                                 ctypes.byref(total), ctypes.byref(lanes), ctypes.byref(plies)))
        return int(total.value), int(lanes.value), int(plies.value)

    def expand_frontier(self, boards_t):
        """One breadth-first ply: returns the SoA tensor of all children (python-chess move order per parent)."""
        n = boards_t.shape[1]
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        check(self.lib.crl_expand_frontier(self.h, _ptr(boards_t), n, None, None, 0, _ptr(counts)))
        offsets = torch.cumsum(counts.to(torch.int64), 0) - counts.to(torch.int64)
        total = int(counts.sum().item())
        out = torch.empty((9, total), dtype=torch.int64, device=self.device)
        check(self.lib.crl_expand_frontier(self.h, _ptr(boards_t), n, _ptr(offsets), _ptr(out), total, _ptr(counts)))
        return out, counts

    def game_replay(self, start_record, moves, records=False):
        """Game semantics for one game on slot 0: returns dict(legal, result, accepted, record[, records]).
        records=True also returns the record after every accepted move ([n_accepted + 1, 9], records[0] = start)."""
        start = np.ascontiguousarray(np.asarray(start_record, dtype=np.uint64))
        mv = np.ascontiguousarray(np.asarray(moves, dtype=np.uint16))
        legal = np.zeros(B.MAX_MOVES, dtype=np.uint16)
        n_legal = ctypes.c_int32(0)
        result = ctypes.c_int8(0)
        accepted = np.zeros(max(len(mv), 1), dtype=np.uint8)
        out = {}
        if records:
            recs = np.zeros((len(mv) + 1, 9), dtype=np.uint64)
            n_rec = ctypes.c_int32(0)
            check(self.lib.crl_game_replay_records_host(
                self.h, _np(start, ctypes.c_uint64), _np(mv, ctypes.c_uint16), len(mv), _np(legal, ctypes.c_uint16),
                ctypes.byref(n_legal), ctypes.byref(result), _np(accepted, ctypes.c_uint8), _np(recs, ctypes.c_uint64),
                ctypes.byref(n_rec)))
            out["records"] = recs[:n_rec.value].copy()
            final = out["records"][-1].copy()
        else:
            final = np.zeros(9, dtype=np.uint64)
            check(self.lib.crl_game_replay_host(self.h, _np(start, ctypes.c_uint64), _np(mv, ctypes.c_uint16), len(mv),
                                                _np(legal, ctypes.c_uint16), ctypes.byref(n_legal), ctypes.byref(result),
                                                _np(accepted, ctypes.c_uint8), _np(final, ctypes.c_uint64)))
        out.update({"legal": legal[:n_legal.value].copy(),
                    "result": None if result.value == B.RESULT_NONE else int(result.value),
                    "accepted": accepted[:len(mv)].astype(bool), "record": final})
        return out

    # ---- encoding -----------------------------------------------------------------------------------------
    def encode(self, boards_t, hist_t=None, hist_len_t=None):
        n = boards_t.shape[1]
        planes = torch.empty((n, 8, 8, 128), dtype=torch.bfloat16, device=self.device)
        check(self.lib.crl_encode(self.h, _ptr(boards_t), _ptr(hist_t), _ptr(hist_len_t), n, _ptr(planes)))
        return planes

    def policy_index(self, moves_t, counts_t):
        n = moves_t.shape[0]
        idx = torch.empty((n, B.MAX_MOVES), dtype=torch.int16, device=self.device)
        check(self.lib.crl_policy_index(self.h, _ptr(moves_t), _ptr(counts_t), n, _ptr(idx)))
        return idx

    def label_table(self):
        t = np.zeros(5 * 4096, dtype=np.int16)
        check(self.lib.crl_label_table_host(self.h, _np(t, ctypes.c_int16)))
        return t.reshape(5, 64, 64)

    # ---- network ------------------------------------------------------------------------------------------
    def load_weights(self, tensors):
        """tensors: the 140-array weight pack (chessrl_b200.model.ChessModel.weights)."""
        arrs = [np.ascontiguousarray(np.asarray(t, dtype=np.float32).reshape(-1)) for t in tensors]
        n = len(arrs)
        ptrs = (_lib.c_f32p * n)(*[a.ctypes.data_as(_lib.c_f32p) for a in arrs])
        sizes = (ctypes.c_int64 * n)(*[a.size for a in arrs])
        check(self.lib.crl_net_load_host(self.h, ptrs, sizes, n))
        self.weights_loaded = True

    def net_forward(self, planes_t):
        n = planes_t.shape[0]
        policy = torch.empty((n, _lib.N_LABELS), dtype=torch.float32, device=self.device)
        value = torch.empty((n,), dtype=torch.float32, device=self.device)
        check(self.lib.crl_net_forward(self.h, _ptr(planes_t), n, _ptr(policy), _ptr(value)))
        return policy, value

    def debug_conv(self, layer, x_t, residual_t=None, relu=False):
        n, cin = x_t.shape[0], x_t.shape[3]
        out = torch.empty((n, 8, 8, 256), dtype=torch.bfloat16, device=self.device)
        check(self.lib.crl_debug_conv(self.h, int(layer), _ptr(x_t), cin, n, _ptr(residual_t), _ptr(out), int(relu)))
        return out

    def debug_tower(self, planes_t, layer=20):
        """The production forward pass with taps: dict(act = output of convolution `layer` [n,8,8,256] bf16,
        logits [n,1968], pf [n,128] bf16, vf [n,64], policy [n,1968], value [n])."""
        n = planes_t.shape[0]
        dev = self.device
        out = {"act": torch.zeros((n, 8, 8, 256), dtype=torch.bfloat16, device=dev),
               "logits": torch.empty((n, _lib.N_LABELS), dtype=torch.float32, device=dev),
               "pf": torch.empty((n, 128), dtype=torch.bfloat16, device=dev),
               "vf": torch.empty((n, 64), dtype=torch.float32, device=dev),
               "policy": torch.empty((n, _lib.N_LABELS), dtype=torch.float32, device=dev),
               "value": torch.empty((n,), dtype=torch.float32, device=dev)}
        check(self.lib.crl_debug_tower(self.h, _ptr(planes_t), n, int(layer), _ptr(out["act"]), _ptr(out["logits"]),
                                       _ptr(out["pf"]), _ptr(out["vf"]), _ptr(out["policy"]), _ptr(out["value"])))
        return out

    def hash_eval(self, boards_t, seed, policy_bits=24):
        n = boards_t.shape[1]
        policy = torch.empty((n, _lib.N_LABELS), dtype=torch.float32, device=self.device)
        value = torch.empty((n,), dtype=torch.float32, device=self.device)
        check(self.lib.crl_hash_eval(self.h, _ptr(boards_t), n, int(seed), int(policy_bits), _ptr(policy), _ptr(value)))
        return policy, value

    # ---- games + search -----------------------------------------------------------------------------------
    def set_evaluator(self, kind, seed=0, policy_bits=24):
        check(self.lib.crl_set_evaluator(self.h, int(kind), int(seed), int(policy_bits)))

    @staticmethod
    def pack_move_lists(move_lists):
        """list of per-game move-word lists -> (uint16 [n, stride], int32 [n]) host arrays for games_set."""
        n = len(move_lists)
        stride = max(1, max((len(m) for m in move_lists), default=1))
        mv = np.full((n, stride), B.MOVE_NONE, dtype=np.uint16)
        cnt = np.zeros(n, dtype=np.int32)
        for i, m in enumerate(move_lists):
            cnt[i] = len(m)
            mv[i, :len(m)] = m
        return mv, cnt

    def games_set(self, start_records, move_lists=None, first=0):
        """Game(board) + Game.move for every listed move.  move_lists: per-game lists, or the packed pair from
        pack_move_lists (host arrays; skips the per-game Python packing)."""
        rec = np.ascontiguousarray(np.asarray(start_records, dtype=np.uint64).reshape(-1, 9))
        n = rec.shape[0]
        if move_lists is None:
            check(self.lib.crl_games_set_host(self.h, first, n, _np(rec, ctypes.c_uint64), None, None, 0))
            return
        if isinstance(move_lists, tuple):
            mv = np.ascontiguousarray(move_lists[0], dtype=np.uint16)
            cnt = np.ascontiguousarray(move_lists[1], dtype=np.int32)
            assert mv.ndim == 2 and mv.shape[0] == n and cnt.shape == (n,)
        else:
            mv, cnt = self.pack_move_lists(move_lists)
        stride = mv.shape[1]
        check(self.lib.crl_games_set_host(self.h, first, n, _np(rec, ctypes.c_uint64), _np(mv, ctypes.c_uint16),
                                          _np(cnt, ctypes.c_int32), stride))

    def games_get(self, first=0, n=None):
        n = self.max_games - first if n is None else n
        rec = np.zeros((n, 9), dtype=np.uint64)
        plies = np.zeros(n, dtype=np.int32)
        res = np.zeros(n, dtype=np.int8)
        check(self.lib.crl_games_get_host(self.h, first, n, _np(rec, ctypes.c_uint64), _np(plies, ctypes.c_int32),
                                          _np(res, ctypes.c_int8)))
        return rec, plies, res

    def games_set_active(self, active, first=0):
        """Parks (0) / resumes (1) lanes first..first+len(active)-1; parked lanes are skipped by the search."""
        a = np.ascontiguousarray(np.asarray(active, dtype=np.uint8))
        check(self.lib.crl_games_set_active_host(self.h, int(first), len(a), _np(a, ctypes.c_uint8)))

    def game_moves(self, game):
        n = ctypes.c_int32(0)
        buf = np.zeros(2048, dtype=np.uint16)
        check(self.lib.crl_game_moves_host(self.h, int(game), _np(buf, ctypes.c_uint16), len(buf), ctypes.byref(n)))
        return buf[:n.value].copy()

    def games_restart(self, lanes, start_record=None):
        """A new game (from `start_record`, default the start position) in every listed lane: one device round trip."""
        ln = np.ascontiguousarray(np.asarray(lanes, dtype=np.int32))
        if ln.size == 0:
            return
        rec = np.ascontiguousarray(B.record_from_fen() if start_record is None else np.asarray(start_record, dtype=np.uint64))
        check(self.lib.crl_games_restart_host(self.h, _np(ln, ctypes.c_int32), int(ln.size), _np(rec, ctypes.c_uint64)))

    def games_moves(self, lanes, cap=2048):
        """Move lists of the listed lanes in one round trip -> list of uint16 arrays."""
        ln = np.ascontiguousarray(np.asarray(lanes, dtype=np.int32))
        if ln.size == 0:
            return []
        buf = np.zeros((ln.size, cap), dtype=np.uint16)
        cnt = np.zeros(ln.size, dtype=np.int32)
        check(self.lib.crl_games_moves_host(self.h, _np(ln, ctypes.c_int32), int(ln.size), _np(buf, ctypes.c_uint16), int(cap),
                                            _np(cnt, ctypes.c_int32)))
        return [buf[i, :min(int(cnt[i]), cap)].copy() for i in range(ln.size)]

    def games_play(self, moves):
        """Game.move of one optional move word per lane (MOVE_NONE = none) -> bool [max_games] accepted."""
        mv = np.ascontiguousarray(np.asarray(moves, dtype=np.uint16))
        assert mv.shape == (self.max_games,)
        acc = np.zeros(self.max_games, dtype=np.uint8)
        check(self.lib.crl_games_play_host(self.h, _np(mv, ctypes.c_uint16), _np(acc, ctypes.c_uint8)))
        torch.cuda.current_stream(self.device).synchronize()
        return acc.astype(bool)

    def games_legal(self, first=0, n=None):
        """Legal moves (python-chess order) of lanes first..first+n-1 -> (uint16 [n, 256], int32 [n])."""
        n = self.max_games - first if n is None else n
        legal = np.zeros((n, B.MAX_MOVES), dtype=np.uint16)
        cnt = np.zeros(n, dtype=np.int32)
        check(self.lib.crl_games_legal_host(self.h, int(first), int(n), _np(legal, ctypes.c_uint16), _np(cnt, ctypes.c_int32)))
        return legal, cnt

    def policy_move(self, mask=None):
        picks = np.zeros(self.max_games, dtype=np.uint16)
        m = None
        if mask is not None:
            mask = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))
            m = _np(mask, ctypes.c_uint8)
        check(self.lib.crl_games_policy_move_host(self.h, m, _np(picks, ctypes.c_uint16)))
        torch.cuda.current_stream(self.device).synchronize()
        return picks

    def mcts_begin_move(self):
        check(self.lib.crl_mcts_begin_move(self.h))

    def set_reuse(self, enable=True):
        """Evaluation reuse across consecutive move searches (include/chessrl_b200.h crl_set_reuse): expansions take the
        reply / value / priors of nodes the previous move's search already evaluated; results are bit-identical."""
        check(self.lib.crl_set_reuse(self.h, int(bool(enable))))
        self.reuse = bool(enable)

    def set_row_bound(self, max_running_games=0):
        """Promise that at most this many games are running (0 = none): launches shrink to that many rows and small batches
        take the tower's single-tile path (include/chessrl_b200.h crl_mcts_set_row_bound).  A broken promise raises."""
        check(self.lib.crl_mcts_set_row_bound(self.h, int(max_running_games)))

    def mcts_simulate(self, n_sims, inflight=1):
        check(self.lib.crl_mcts_simulate(self.h, int(n_sims), int(inflight)))

    def root_stats(self, want=("visits", "values", "priors", "moves", "replies", "results")):
        G, M = self.max_games, B.MAX_MOVES
        out = {
            "visits": np.zeros((G, M), dtype=np.int32) if "visits" in want else None,
            "values": np.zeros((G, M), dtype=np.float64) if "values" in want else None,
            "priors": np.zeros((G, M), dtype=np.float32) if "priors" in want else None,
            "moves": np.zeros((G, M), dtype=np.uint16) if "moves" in want else None,
            "replies": np.zeros((G, M), dtype=np.uint16) if "replies" in want else None,
            "results": np.zeros((G, M), dtype=np.int8) if "results" in want else None,
            "n_children": np.zeros(G, dtype=np.int32),
            "root_visits": np.zeros(G, dtype=np.int32),
            "root_values": np.zeros(G, dtype=np.float64),
        }

        def p(name, ct):
            return _np(out[name], ct) if out[name] is not None else None

        check(self.lib.crl_mcts_root_stats_host(self.h, p("visits", ctypes.c_int32), p("values", ctypes.c_double),
                                                p("priors", ctypes.c_float), p("moves", ctypes.c_uint16),
                                                p("replies", ctypes.c_uint16), p("results", ctypes.c_int8),
                                                p("n_children", ctypes.c_int32), p("root_visits", ctypes.c_int32),
                                                p("root_values", ctypes.c_double)))
        return out

    def commit(self, picks, apply=True):
        pk = np.ascontiguousarray(np.asarray(picks, dtype=np.int32))
        out = np.zeros((self.max_games, 2), dtype=np.uint16)
        check(self.lib.crl_mcts_commit_host(self.h, _np(pk, ctypes.c_int32), _np(out, ctypes.c_uint16), int(bool(apply))))
        torch.cuda.current_stream(self.device).synchronize()
        return out

    def node_dump(self, game=0, cap=None):
        cap = self.max_nodes + 1 if cap is None else cap
        arr = (_lib.NodeHost * cap)()
        n = ctypes.c_int32(0)
        check(self.lib.crl_mcts_node_dump_host(self.h, int(game), arr, cap, ctypes.byref(n)))
        return [arr[i] for i in range(min(n.value, cap))]

    def counters(self):
        c = np.zeros(3, dtype=np.int64)
        check(self.lib.crl_counters_host(self.h, _np(c, ctypes.c_int64)))
        r = np.zeros(1, dtype=np.int64)
        check(self.lib.crl_reuse_count_host(self.h, _np(r, ctypes.c_int64)))
        return {"simulations": int(c[0]), "evaluations": int(c[1]), "launches": int(c[2]), "reused_evaluations": int(r[0])}

    def profile(self, enable):
        check(self.lib.crl_profile(self.h, int(bool(enable))))

    def profile_read(self):
        k = len(_lib.KERNEL_CLASSES)
        ms = np.zeros(k, dtype=np.float64)
        ln = np.zeros(k, dtype=np.int64)
        check(self.lib.crl_profile_read_host(self.h, _np(ms, ctypes.c_double), _np(ln, ctypes.c_int64), k))
        return {name: {"ms": float(ms[i]), "launches": int(ln[i])} for i, name in enumerate(_lib.KERNEL_CLASSES)}
