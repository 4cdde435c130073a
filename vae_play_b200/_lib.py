"""ctypes binding of libvaeplay_b200.so (include/vaeplay_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that is impossible the
import raises.  Every compute entry point takes raw device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvaeplay_b200.so")
_EXPERIMENT_LIB = os.environ.get("VP_LIB_PATH")       # A/B experiments only: load another build of the same sources

F32, BF16 = 0, 1
ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2, "tanh": 3, "sigmoid": 4}
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
ABI_VERSION = 2


class VpConvGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n", "hi", "wi", "ci", "ho", "wo", "co", "kh", "kw", "stride", "pad", "transposed")]


class VaePlayError(RuntimeError):
    pass


_i = C.c_int
_u64 = C.c_uint64

HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vaeplay_b200.h")

_CTYPE = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "size_t": C.c_size_t,
          "float": C.c_float, "double": C.c_double, "unsigned int": C.c_uint, "unsigned": C.c_uint}


def _arg_ctype(decl: str):
    """ctypes type of one C parameter declaration of include/vaeplay_b200.h."""
    decl = decl.strip()
    if "*" in decl:
        base = decl[: decl.index("*")].replace("const", "").strip()
        if base == "VpConvGeom":
            return C.POINTER(VpConvGeom)
        if base == "int" and decl.count("*") == 1:
            return C.POINTER(C.c_int)          # host out-parameter (int* nparts)
        return C.c_void_p                      # device / host arrays passed as raw addresses
    words = decl.replace("const", "").split()
    base = " ".join(words[:-1]) if len(words) > 1 else words[0]
    if base not in _CTYPE:
        raise VaePlayError(f"include/vaeplay_b200.h: unknown parameter type in '{decl}'")
    return _CTYPE[base]


def parse_header(path: str = HEADER_PATH):
    """name -> (restype, [argtypes]) for every prototype the header declares: the ctypes table is GENERATED from the
    header, so an ABI change cannot drift silently past the binding."""
    import re
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    out = {}
    for m in re.finditer(r"(const\s+char\s*\*|uint64_t|int)\s+(vp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        res = C.c_char_p if "char" in ret else (_u64 if ret == "uint64_t" else _i)
        out[name] = (res, [] if args in ("", "void") else [_arg_ctype(a) for a in args.split(",")])
    return out


_PROTOS = parse_header()
PLAIN = {k: v for k, v in _PROTOS.items() if k in ("vp_last_error", "vp_abi_version", "vp_device_arch", "vp_launch_count", "vp_simt_bf16_count")}
SIGNATURES = {k: v[1] for k, v in _PROTOS.items() if k not in PLAIN}     # compute entry points: int return code

_lib = None


def load(build_if_missing: bool = True):
    """Load the CUDA library, (re)building it first when it is missing or older than its sources (content hash, see
    build.sources_digest).  Raises if it cannot be had: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _EXPERIMENT_LIB:
        pass
    elif not os.path.exists(LIB_PATH) or _build.is_stale():
        if not build_if_missing and not os.path.exists(LIB_PATH):
            raise VaePlayError(f"{LIB_PATH} is missing; run `python -m vae_play_b200.build`")
        try:
            _build.build()
        except Exception as e:       # no nvcc on this machine, ...
            if not os.path.exists(LIB_PATH):
                raise VaePlayError(f"cannot build {LIB_PATH}: {e}") from e
            raise VaePlayError(f"{LIB_PATH} is older than its sources and cannot be rebuilt: {e}") from e
    lib = C.CDLL(_EXPERIMENT_LIB or LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    for name, (res, args) in PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if lib.vp_abi_version() != ABI_VERSION:
        raise VaePlayError("libvaeplay_b200.so ABI version mismatch; rebuild with `python -m vae_play_b200.build --force`")
    _lib = lib
    return lib


_workspace = {}


def ensure_workspace(device, nbytes: int = 32 << 20):
    """Register a per-device scratch buffer for split-K partial sums (vp_set_workspace keeps one pointer per device)."""
    import torch
    key = (device.type, device.index)
    if key not in _workspace:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspace[key] = buf
        call("vp_set_workspace", C.c_void_p(buf.data_ptr()), nbytes)
    return _workspace[key]


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vp_last_error()
        raise VaePlayError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)


def launch_count() -> int:
    return int(load().vp_launch_count())


def simt_bf16_count() -> int:
    return int(load().vp_simt_bf16_count())
