"""netencoder: board -> network planes and move -> policy index, drop-in for the reference's netencoder.py.

get_game_state(game, flipped=False) -> float64 [8, 8, 127] (netencoder.py:72-91) is produced by the CUDA encode
kernel (crl_encode) from the game's current record and its last 8 previous positions; get_uci_labels()
(netencoder.py:94-134) returns the 1968 move labels whose order defines the policy head.  DataGameSequence
(netencoder.py:137-181) keeps its interface; it is the "next" row of the scope table (training input pipeline).
"""

from __future__ import annotations

import numpy as np

from . import boards as B
from . import runtime

_FILES = "abcdefgh"


def get_game_state(game, flipped=False):
    import torch
    eng = runtime.scalar_engine()
    recs = game.history_records()
    boards = eng.boards_to_device(np.asarray(recs[0], dtype=np.uint64)[None, :])
    hist = np.zeros((8, 8, 1), dtype=np.uint64)
    for i, r in enumerate(recs[1:9]):
        hist[i, :, 0] = r[:8]
    hist_t = torch.from_numpy(hist.view(np.int64)).to(eng.device)
    hlen_t = torch.tensor([len(recs) - 1], dtype=torch.uint8, device=eng.device)
    planes = eng.encode(boards, hist_t, hlen_t)
    out = planes[0, :, :, :127].to(torch.float64).cpu().numpy()
    if flipped:
        out = np.rot90(out, k=2)
    return out


def get_uci_labels():
    """The 1968 UCI move labels in policy-head order: for every source square (file-major a1, a2, ... h8) the
    queen-line and knight destinations, then the under/promotions per file (netencoder.py:94-134)."""
    labels = []
    knight = ((-2, -1), (-1, -2), (-2, 1), (1, -2), (2, -1), (-1, 2), (2, 1), (1, 2))
    for f in range(8):
        for r in range(8):
            targets = [(k, r) for k in range(8)] + [(f, k) for k in range(8)]
            targets += [(f + k, r + k) for k in range(-7, 8)] + [(f + k, r - k) for k in range(-7, 8)]
            targets += [(f + a, r + b) for a, b in knight]
            src = _FILES[f] + str(r + 1)
            labels += [src + _FILES[tf] + str(tr + 1) for tf, tr in targets
                       if (tf, tr) != (f, r) and 0 <= tf < 8 and 0 <= tr < 8]
    for f in range(8):
        for piece in "qrbn":
            for df in (0, -1, 1):
                if 0 <= f + df < 8:
                    labels.append("%s2%s1%s" % (_FILES[f], _FILES[f + df], piece))
                    labels.append("%s7%s8%s" % (_FILES[f], _FILES[f + df], piece))
    return labels


class DataGameSequence(object):
    """Batches of (planes, (policy one-hot, value)) from a DatasetGame; one sample per ply of `batch_size` games
    (netencoder.py:137-181).  random_flips = probability of rotating a whole game's planes by 180 degrees."""

    def __init__(self, dataset, batch_size=8, random_flips=0):
        self.dataset = dataset
        self.batch_size = min(batch_size, len(dataset))
        self.uci_ids = {u: i for i, u in enumerate(get_uci_labels())}
        self.random_flips = random_flips

    def __len__(self):
        return int(len(self.dataset) / self.batch_size)

    def __getitem__(self, idx):
        """(planes float64 [N,8,8,127], (policy one-hot float32 [N,1968], value [N])), N = plies of the batch's games.
        All plies are encoded in one crl_encode launch (training.encode_games) instead of one get_game_state call
        per sample; the flip draw is the reference's (one np.random.rand() per game)."""
        import torch
        from . import training
        games = self.dataset[idx * self.batch_size:(idx + 1) * self.batch_size]
        flips = training.draw_flips(len(games), self.random_flips)
        planes, pol, _ = training.encode_games(games, flips)
        xs = planes[..., :127].to(torch.float64).cpu().numpy()
        pol = pol.cpu().numpy()
        onehot = np.zeros((len(pol), 1968), dtype=np.float32)
        onehot[np.arange(len(pol)), pol] = 1.0
        vals = []
        for g in games:
            vals.extend([g.get_result()] * len(g))               # the result as stored: None for an unfinished game
        return xs, (onehot, np.asarray(vals))
