"""ctypes binding of libxmc.so. Prototypes are taken from include/xmc.h (the single source of truth of the C ABI).
The product path has no CPU fallback: if the library is missing or a call fails, an exception is raised."""
import ctypes
import os
import re

import torch

from . import build as _build

_LIB = None
HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "xmc.h")


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
      ("strideH", ctypes.c_int), ("strideW", ctypes.c_int), ("Hin", ctypes.c_int), ("Win", ctypes.c_int),
      ("pitchW", ctypes.c_longlong), ("pitchH", ctypes.c_longlong), ("pitchN", ctypes.c_longlong),
      ("mask_last", ctypes.c_int), ("subpixel", ctypes.c_int), ("act_f32", ctypes.c_int), ("ldPair", ctypes.c_int),
      ("res_pair", ctypes.c_int), ("mask_pair", ctypes.c_int),
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
      ("alpha", ctypes.c_float), ("subpixel", ctypes.c_int),
      ("HinA", ctypes.c_int),
      ("pitchWA", ctypes.c_longlong), ("pitchHA", ctypes.c_longlong), ("pitchNA", ctypes.c_longlong),
  ]


class BnDesc(ctypes.Structure):
  _fields_ = [
      ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("C", ctypes.c_int),
      ("Hc", ctypes.c_int),
      ("ldG", ctypes.c_int), ("goff", ctypes.c_int), ("boff", ctypes.c_int),
      ("relu", ctypes.c_int), ("upsample", ctypes.c_int), ("replicas", ctypes.c_int), ("act_f32", ctypes.c_int),
  ]


class PrepEntry(ctypes.Structure):
  _fields_ = [
      ("w_off", ctypes.c_longlong), ("wk_fwd_off", ctypes.c_longlong), ("wk_dg_off", ctypes.c_longlong),
      ("bias_off", ctypes.c_longlong), ("bias_dst_off", ctypes.c_longlong),
      ("taps", ctypes.c_int), ("cin", ctypes.c_int), ("cout", ctypes.c_int),
      ("ld_fwd", ctypes.c_int), ("ld_dg", ctypes.c_int),
      ("sn", ctypes.c_int), ("tile_begin", ctypes.c_int), ("split", ctypes.c_int),
      ("cscale_off", ctypes.c_longlong), ("dg_part_stride", ctypes.c_longlong),
  ]


class SnEntry(ctypes.Structure):
  _fields_ = [
      ("w_off", ctypes.c_longlong), ("t_off", ctypes.c_longlong), ("s_off", ctypes.c_longlong),
      ("u_off", ctypes.c_longlong),
      ("rows", ctypes.c_int), ("cols", ctypes.c_int),
      ("row_block_begin", ctypes.c_int), ("col_tile_begin", ctypes.c_int), ("elem_block_begin", ctypes.c_int),
      ("reserved", ctypes.c_int),
  ]


_CTYPE = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double}


def declared_functions(header=HEADER):
  """Parses include/xmc.h and returns {name: (restype, [argtypes])}."""
  src = open(header).read()
  src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
  out = {}
  for m in re.finditer(r"\b(int|const char\*)\s+(xmc_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
    ret, name, args = m.group(1), m.group(2), m.group(3)
    argtypes = []
    args = args.strip()
    if args and args != "void":
      for a in args.split(","):
        a = a.strip()
        if "*" in a:
          argtypes.append(ctypes.c_void_p)
        else:
          t = re.sub(r"\b\w+$", "", a).strip()  # drop the parameter name
          argtypes.append(_CTYPE[t])
    out[name] = (ctypes.c_char_p if "char" in ret else ctypes.c_int, argtypes)
  return out


def lib():
  """Loads (building first if the in-tree .so is missing) and returns the ctypes handle."""
  global _LIB
  if _LIB is not None:
    return _LIB
  path = _build.LIB_PATH
  if not os.path.exists(path):
    path = _build.build()
  L = ctypes.CDLL(path)
  for name, (res, args) in declared_functions().items():
    fn = getattr(L, name)  # AttributeError here means the .so does not export a declared symbol
    fn.restype = res
    fn.argtypes = args
  # the ctypes mirrors above must have the layout the library was compiled with
  for which, cls in enumerate((ConvDesc, WgradDesc, BnDesc, PrepEntry, SnEntry)):
    if L.xmc_sizeof(which) != ctypes.sizeof(cls):
      raise XmcError(f"ABI mismatch: sizeof({cls.__name__}) = {ctypes.sizeof(cls)}, library says {L.xmc_sizeof(which)}")
  _LIB = L
  return L


def check(code):
  if code != 0:
    L = lib()
    raise XmcError(f"libxmc: {L.xmc_strerror(code).decode()} [{L.xmc_last_cuda_error().decode()}]")


def ptr(t):
  if t is None:
    return None
  return t.data_ptr()


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
  """Raw cudaStream_t of torch's current stream on the current device (every kernel of the library launches on it).
  The raw getter is ~10x cheaper than building a torch.cuda.Stream object, and this runs once per kernel launch."""
  if _RAW_STREAM is not None:
    return _RAW_STREAM(torch.cuda.current_device())
  return torch.cuda.current_stream().cuda_stream
