"""Logger set-up with the reference's verbosity levels (scGeneClust/_utils.py:77-98).  The dataset loaders of the
reference (network downloads through scanpy/squidpy) are out of scope."""
import sys

from loguru import logger


def set_logger(verbosity: int = 1):
    """verbosity 0 / 1 / 2 -> WARNING / INFO / DEBUG on stdout."""
    def formatter(record: dict):
        if record['level'].name in ('DEBUG', 'INFO'):
            return "<level>{level: <5}</level> | <cyan>{name}</cyan>:<cyan>{function}</cyan>:<cyan>{line}</cyan> - " \
                   "<level>{message}</level>\n{exception}"
        return "<level>{level: <8}</level> | <cyan>{name}</cyan>:<cyan>{function}</cyan>:<cyan>{line}</cyan> - " \
               "<level>{message}</level>\n{exception}"
    global _CONFIGURED
    level = {0: 'WARNING', 1: 'INFO', 2: 'DEBUG'}[verbosity]
    if _CONFIGURED == (level, id(sys.stdout)):
        return                          # re-installing the same sink costs ~4 ms (loguru rebuilds its exception formatter)
    logger.remove()
    logger.add(sys.stdout, colorize=True, level=level, format=formatter)
    _CONFIGURED = (level, id(sys.stdout))


_CONFIGURED = None
