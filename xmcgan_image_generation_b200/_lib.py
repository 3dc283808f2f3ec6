"""ctypes binding of libxmc.so (declared in include/xmc.h). The product path has no CPU fallback: if the library is
missing or a call fails, an exception is raised."""
import ctypes
import os

import torch

from . import build as _build

_LIB = None


class XmcError(RuntimeError):
  pass


class ConvDesc(ctypes.Structure):
  _fields_ = [
      ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("C", ctypes.c_int),
      ("ldA", ctypes.c_int),
      ("KH", ctypes.c_int), ("KW", ctypes.c_int), ("pad_h", ctypes.c_int), ("pad_w", ctypes.c_int),
      ("Cout", ctypes.c_int),
      ("ldB", ctypes.c_int),
      ("strideB_batch", ctypes.c_longlong),
      ("batched", ctypes.c_int),
      ("out_dtype", ctypes.c_int),
      ("ldOut", ctypes.c_int),
      ("alpha", ctypes.c_float),
      ("relu", ctypes.c_int),
      ("res_shift", ctypes.c_int),
      ("ldRes", ctypes.c_int), ("ldMask", ctypes.c_int),
  ]


class WgradDesc(ctypes.Structure):
  _fields_ = [
      ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
      ("Ca", ctypes.c_int), ("ldA", ctypes.c_int),
      ("Cb", ctypes.c_int), ("ldB", ctypes.c_int),
      ("KH", ctypes.c_int), ("KW", ctypes.c_int), ("pad_h", ctypes.c_int), ("pad_w", ctypes.c_int),
      ("batched", ctypes.c_int),
      ("out_mode", ctypes.c_int),
      ("ldOut", ctypes.c_int),
      ("out_tap_stride", ctypes.c_longlong), ("out_batch_stride", ctypes.c_longlong),
      ("alpha", ctypes.c_float),
  ]


def lib():
  """Loads (building first if the in-tree .so is stale and nvcc is available) and returns the ctypes handle."""
  global _LIB
  if _LIB is not None:
    return _LIB
  path = _build.LIB_PATH
  if not os.path.exists(path):
    path = _build.build()
  L = ctypes.CDLL(path)
  L.xmc_strerror.restype = ctypes.c_char_p
  L.xmc_last_cuda_error.restype = ctypes.c_char_p
  _LIB = L
  return L


def check(code):
  if code != 0:
    L = lib()
    raise XmcError(f"libxmc: {L.xmc_strerror(code).decode()} [{L.xmc_last_cuda_error().decode()}]")


def ptr(t):
  if t is None:
    return ctypes.c_void_p(0)
  return ctypes.c_void_p(t.data_ptr())


def stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
