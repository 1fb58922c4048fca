"""chessrl_b200: B200-native lockstep self-play engine behind the Python surface of AIRLegend/ChessRL.

Submodules mirror the reference's flat modules (game, agent, agentdistributed, mctree, netencoder, dataset, selfplay,
supervised); every computation happens in libchessrl_b200.so (include/chessrl_b200.h) on an sm_100a device.  Nothing is
imported here, so `import chessrl_b200` works without a GPU; the engine raises as soon as it is asked to compute."""

__version__ = "0.1.0"
