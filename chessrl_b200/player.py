"""Player: the abstract base every chess-playing object derives from (reference player.py:1-14).

Contract kept from the reference: a bare Player cannot be instantiated (Exception), subclasses carry `color`
(True = white) and answer best_move(game) with a UCI string.
"""


class Player(object):

    color = None

    def __init__(self, color):
        self._refuse_bare_base()
        self.color = color

    def _refuse_bare_base(self):
        if self.__class__ is Player:
            raise Exception("Player is abstract: use Agent, AgentDistributed or another subclass")

    def best_move(self, game) -> str:
        """Subclasses return the move to play in `game` (UCI)."""
        raise Exception("best_move() is not implemented by %s" % self.__class__.__name__)
