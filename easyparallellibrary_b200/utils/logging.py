"""Rank-aware logging (the reference logs through ``tf_logging`` only)."""
import logging
import os

_LOGGER = None


def get_logger() -> logging.Logger:
  global _LOGGER
  if _LOGGER is None:
    logger = logging.getLogger("epl_b200")
    if not logger.handlers:
      h = logging.StreamHandler()
      h.setFormatter(logging.Formatter("[epl r%s %%(levelname)s %%(asctime)s] %%(message)s" % os.environ.get("RANK", "0"),
                                       "%H:%M:%S"))
      logger.addHandler(h)
    logger.setLevel(os.environ.get("EPL_LOG_LEVEL", "WARNING").upper())
    logger.propagate = False
    _LOGGER = logger
  return _LOGGER


def rank0_info(msg, *args):
  if os.environ.get("RANK", "0") == "0":
    get_logger().info(msg, *args)
