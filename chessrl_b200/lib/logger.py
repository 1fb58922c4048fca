"""Logger: singleton wrapper with the reference's levels 0 (debug) / 1 (info) / 2 (warning) (lib/logger.py)."""
import logging


class Logger(object):
    _instance = None

    def __init__(self):
        self._log = logging.getLogger("chessrl_b200")
        if not self._log.handlers:
            h = logging.StreamHandler()
            h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s"))
            self._log.addHandler(h)
        self.set_level(1)

    @classmethod
    def get_instance(cls):
        if cls._instance is None:
            cls._instance = Logger()
        return cls._instance

    def set_level(self, level):
        self._log.setLevel({0: logging.DEBUG, 1: logging.INFO, 2: logging.WARNING}.get(level, logging.INFO))

    def debug(self, msg):
        self._log.debug(msg)

    def info(self, msg):
        self._log.info(msg)

    def warning(self, msg):
        self._log.warning(msg)

    def error(self, msg):
        self._log.error(msg)
