"""Player: abstract base of every chess player object (reference player.py)."""


class Player(object):

    def __init__(self, color):
        if type(self) is Player:
            raise Exception('Cannot create Player Abstract class.')
        self.color = color

    def best_move(self, game) -> str:
        raise Exception('Abstract class.')
