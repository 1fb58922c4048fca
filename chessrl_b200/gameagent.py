"""GameAgent: drop-in for the reference's game-versus-agent environment (gameagent.py:8-51).

A Game in which an Agent answers every accepted move with the argmax of its legal-masked policy
(Agent.best_move(real_game=True), agent.py:39-43); when the agent plays white it opens on the first call to move().
Same constructor contract: `agent` is an Agent or a path to its weights, anything else raises ValueError
(gameagent.py:18-23).  All rules and the network evaluation run in the CUDA engine through Game / Agent.
"""

from __future__ import annotations

from .agent import Agent
from .game import Game


def _as_agent(agent, player_color):
    if isinstance(agent, Agent):
        return agent
    if type(agent) == str:                       # a weights path: the agent takes the colour the player did not
        return Agent(not player_color, weights=agent)
    raise ValueError("An agent or path to the agents weights (.h5) is needed")


class GameAgent(Game):

    def __init__(self, agent, player_color=Game.WHITE, board=None, date=None):
        super().__init__(board=board, player_color=player_color, date=date)
        self.agent = _as_agent(agent, player_color)

    def _agent_plays(self):
        return Game.move(self, self.agent.best_move(self, real_game=True))

    def move(self, movement):
        """Plays `movement` and lets the agent reply; an illegal move is ignored and returns False.  On an empty
        board with the agent as white, the agent opens instead and `movement` is not played (gameagent.py:25-43)."""
        if self.agent.color and not self.board.move_stack:
            return self._agent_plays()
        accepted = Game.move(self, movement)
        if accepted and self.get_result() is None:
            self._agent_plays()
        return accepted

    def get_copy(self):
        twin = Game.get_copy(self)
        twin.__class__ = GameAgent
        twin.agent, twin.player_color = self.agent, self.player_color
        return twin

    def tearup(self):
        """Free resources."""
        del self.agent
