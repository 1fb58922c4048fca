"""DatasetGame: list of games with the reference's JSON format (dataset.py:6-97):
[{"moves": [uci...], "result": 1|0|-1|null, "player_color": bool, "date": "dd/mm/YYYY HH:MM:SS"}, ...]."""

from __future__ import annotations

import json

from . import game


class DatasetGame(object):

    def __init__(self, games=None):
        self.games = games if games is not None else []

    def augment_game(self, game_base):
        """One {'game', 'next_move', 'result'} sample per ply of the game (dataset.py:21-43)."""
        hist = game_base.get_history()
        g = game.Game(date=hist['date'], player_color=hist['player_color'])
        out = []
        for m in hist['moves']:
            out.append({'game': g, 'next_move': m, 'result': hist['result']})
            g = g.get_copy()
            g.move(m)
        return out

    def load(self, path):
        with open(path, 'r') as f:
            self.loads(f.read())

    def loads(self, string):
        for item in json.loads(string):
            if len(item['moves']) > 0:
                g = game.Game(date=item['date'], player_color=item['player_color'])
                g._sync(extra=item['moves'])
                self.games.append(g)

    def save(self, path):
        """Appends to what the file already holds (dataset.py:60-71)."""
        existing = DatasetGame()
        try:
            existing.load(path)
        except FileNotFoundError:
            pass
        with open(path, 'w') as f:
            json.dump([g.get_history() for g in existing.games + self.games], f)

    def append(self, other):
        if isinstance(other, game.Game):
            self.games.append(other)
        elif isinstance(other, DatasetGame):
            self.games.extend(other.games)

    def __str__(self):
        return json.dumps([g.get_history() for g in self.games])

    def __add__(self, other):
        self.append(other)
        return self

    def __iadd__(self, other):
        return self.__add__(other)

    def __len__(self):
        return len(self.games)

    def __getitem__(self, key):
        return self.games[key]
