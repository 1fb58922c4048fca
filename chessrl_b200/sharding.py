"""Multi-GPU plumbing: games are independent, so lanes shard by rank with NO data-path collective.
torch.distributed is used only to (i) broadcast the weight pack from rank 0 and (ii) gather finished games
(moves int16 + length + result + colour) on rank 0.  Backend: NCCL on GPUs, gloo in the CPU tests."""

from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import boards as B

MAX_PLIES = 2048


def init(backend=None):
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    dist.init_process_group(backend)


def shutdown():
    """Leave the process group (a run that exits without doing so gets a resource-leak warning from NCCL)."""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def shard_range(n_items, rank=None, world=None):
    """Contiguous block of items owned by `rank` (first ranks take the remainder)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sum_counts(counts):
    """perft across ranks: every rank holds the node counts of ITS frontier boards (board i -> rank i % world);
    one all_reduce(sum) of an int64 joins them.  Returns the total as a Python int on every rank."""
    t = torch.as_tensor(counts).to(torch.int64).sum().reshape(1).to(_dev())
    dist.all_reduce(t)
    return int(t.item())


def broadcast_weights(model, src=0):
    flat = torch.cat([torch.from_numpy(np.ascontiguousarray(w).reshape(-1)) for w in model.weights]).to(_dev())
    dist.broadcast(flat, src)
    flat = flat.cpu().numpy()
    out, o = [], 0
    for w in model.weights:
        out.append(flat[o:o + w.size].reshape(w.shape).astype(np.float32))
        o += w.size
    model.weights = out


def pack_games(move_words, results, colors):
    """-> (moves int16 [n, L], lengths int32 [n], results int8 [n], colors uint8 [n])"""
    n = len(move_words)
    L = max([len(m) for m in move_words] + [1])
    mv = np.full((n, L), -1, dtype=np.int16)
    ln = np.zeros(n, dtype=np.int32)
    for i, m in enumerate(move_words):
        ln[i] = len(m)
        mv[i, :len(m)] = np.asarray(m, dtype=np.uint16).view(np.int16)
    res = np.array([B.RESULT_NONE if r is None else r for r in results], dtype=np.int8)
    return mv, ln, res, np.asarray(colors, dtype=np.uint8)


def gather_packed(mv, ln, res, col, dst=0):
    """All ranks call; rank `dst` gets the concatenation in rank order, others get None."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _dev()
    meta = torch.tensor([mv.shape[0], mv.shape[1]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    n_max = int(max(m[0].item() for m in metas))
    l_max = int(max(m[1].item() for m in metas))
    pad = torch.full((n_max, l_max + 3), -1, dtype=torch.int32, device=dev)   # gloo has no int16
    if mv.shape[0]:
        pad[:mv.shape[0], :mv.shape[1]] = torch.from_numpy(mv.astype(np.int32)).to(dev)
        pad[:mv.shape[0], l_max] = torch.from_numpy(ln.astype(np.int32)).to(dev)
        pad[:mv.shape[0], l_max + 1] = torch.from_numpy(res.astype(np.int32)).to(dev)
        pad[:mv.shape[0], l_max + 2] = torch.from_numpy(col.astype(np.int32)).to(dev)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    out = []
    for m, b in zip(metas, bufs):
        b = b[:int(m[0].item())].cpu().numpy()
        for row in b:
            n = int(row[l_max])
            out.append((row[:n].astype(np.int16).view(np.uint16).copy(), int(row[l_max + 1]), bool(row[l_max + 2])))
    return out


def gather_games(dataset, dst=0):
    """DatasetGame on every rank -> merged DatasetGame on rank dst (empty elsewhere)."""
    from .dataset import DatasetGame
    from .game import Game
    words = [[B.uci_to_move(m) for m in g._moves] for g in dataset.games]
    packed = pack_games(words, [g.get_result() for g in dataset.games], [g.player_color for g in dataset.games])
    got = gather_packed(*packed, dst=dst)
    merged = DatasetGame()
    if got is not None:
        for mv, res, col in got:
            g = Game(player_color=col)
            g._sync(extra=[B.move_to_uci(m) for m in mv])
            merged.append(g)
    return merged
