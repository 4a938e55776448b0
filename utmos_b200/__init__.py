"""utmos_b200 -- B200-native greedy maximum-coverage sample selection (drop-in for `utmos select/convert`).

Host side is Python mirroring utmos/select.py and utmos/convert.py of ACEnglish/utmos v2.2.0; all numeric
work happens in hand-written sm_100a CUDA behind the C ABI in include/utmos_b200.h (no CPU fallback).
"""
__version__ = "2.2.0"          # the reference version this package is a drop-in for (utmos/__init__.py:5)
__b200_version__ = "0.1.0"
