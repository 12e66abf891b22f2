"""screencounter_b200: B200-native barcode counting behind screenCounter's entry points.

Layers:
  csrc/      hand-written sm_100a kernels + C++ host layer behind the C ABI of include/scg.h
  rcpp.py    the reference's seven Rcpp-level functions (same names, argument order, return shapes)
  device.py  resident reads / plans for callers that keep data in HBM (bench.py)
  multi.py   one process per GPU: dense counts by one all-reduce, sparse tables merged on the devices
"""
from . import rcpp  # noqa: F401
from .rcpp import ScreenCounterError  # noqa: F401

__version__ = "0.1.0"
