"""GameAgent: drop-in for the reference's game-versus-agent environment (gameagent.py:8-51).

A Game in which an Agent answers every accepted move with the argmax of its legal-masked policy
(Agent.best_move(real_game=True), agent.py:39-43); when the agent plays white it opens on the first call to move().
Same constructor contract: `agent` is an Agent or a path to its weights, anything else raises ValueError
(gameagent.py:18-23).  All rules and the network evaluation run in the CUDA engine through Game / Agent.
"""

from __future__ import annotations

from .agent import Agent
from .game import Game


class GameAgent(Game):

    def __init__(self, agent, player_color=Game.WHITE, board=None, date=None):
        super().__init__(board=board, player_color=player_color, date=date)
        if isinstance(agent, Agent):
            self.agent = agent
        elif type(agent) == str:
            self.agent = Agent(not player_color, weights=agent)
        else:
            raise ValueError("An agent or path to the agents weights (.h5) is needed")

    def move(self, movement):
        """Makes a move; the agent replies.  An illegal move is ignored (returns False) (gameagent.py:25-43)."""
        if self.agent.color and len(self.board.move_stack) == 0:
            # agent plays white and nothing has been played yet: it opens, `movement` is not played
            return super().move(self.agent.best_move(self, real_game=True))
        made_movement = super().move(movement)
        if made_movement and self.get_result() is None:
            super().move(self.agent.best_move(self, real_game=True))
        return made_movement

    def get_copy(self):
        g = Game.get_copy(self)
        g.__class__ = GameAgent
        g.agent = self.agent
        g.player_color = self.player_color
        return g

    def tearup(self):
        """Free resources."""
        del self.agent
