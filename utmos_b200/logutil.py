"""Logging setup equivalent to the ``truvari.setup_logging`` call sites (utmos/select.py:400, utmos/convert.py:38)."""
import logging
import sys


def setup_logging(debug=False, stream=sys.stderr):
    """stderr logging at INFO (DEBUG with --debug)."""
    level = logging.DEBUG if debug else logging.INFO
    logging.basicConfig(stream=stream, level=level, format="%(asctime)s [%(levelname)s] %(message)s", force=True)
