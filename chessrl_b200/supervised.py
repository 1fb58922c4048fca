"""supervised: drop-in for the reference's entry point (supervised.py:66-95)

    python -m chessrl_b200.supervised modeldir datadir [--epochs 1] [--bs 8] [--debug]
"""

from __future__ import annotations

import argparse

from .agent import Agent
from .dataset import DatasetGame
from .lib.logger import Logger
from .selfplay import get_model_path


def train(model_dir, dataset_path, epochs=1, batch_size=8):
    logger = Logger.get_instance()
    logger.info("Loading dataset")
    data_train = DatasetGame()
    data_train.load(dataset_path)
    model_path = get_model_path(model_dir)
    logger.info("Loading the agent...")
    chess_agent = Agent(color=True)
    try:
        chess_agent.load(model_path)
    except OSError:
        logger.warning("Model not found, training a fresh one.")
    chess_agent.train(data_train, logdir=model_dir, epochs=epochs, validation_split=0.25, batch_size=batch_size)
    logger.info("Saving the agent...")
    chess_agent.save(model_path)


def main(argv=None):
    parser = argparse.ArgumentParser(description="Trains a model on a dataset of recorded games.")
    parser.add_argument('model_dir', metavar='modeldir', help="where to store (and load from) the trained model and the logs")
    parser.add_argument('data_path', metavar='datadir', help="Path of .JSON dataset.")
    parser.add_argument('--epochs', metavar='epochs', type=int, default=1)
    parser.add_argument('--bs', metavar='bs', help="Batch size. Default 8", type=int, default=8)
    parser.add_argument('--debug', action='store_true', default=False, help="Log debug messages on screen. Default false.")
    args = parser.parse_args(argv)
    logger = Logger.get_instance()
    logger.set_level(0 if args.debug else 1)
    train(args.model_dir, args.data_path, args.epochs, args.bs)


if __name__ == "__main__":
    main()
