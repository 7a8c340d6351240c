"""ctypes binding of the 15-function ``mpeg1_decoder_*`` C ABI (reference src/wasm/mpeg1.h:10-25).

The same binding works for any shared library that exports that ABI: the product
(``jsmpeg_b200/libjsmpeg_b200.so`` -- CUDA) and, in tests only, the CPU checkers the test suite
builds (the compiled reference and our restatement of it).  Nothing in this package loads those.
"""
from __future__ import annotations

import ctypes
import os

BIT_BUFFER_MODE_EVICT = 1   # reference src/wasm/buffer.h:8-11
BIT_BUFFER_MODE_EXPAND = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.environ.get("JSMPEG_B200_LIB", os.path.join(_HERE, "libjsmpeg_b200.so"))  # override: kernel-tuning experiments

MPEG1_ABI = {
    # name: (restype, argtypes)
    "mpeg1_decoder_create": (ctypes.c_void_p, [ctypes.c_uint, ctypes.c_int]),
    "mpeg1_decoder_destroy": (None, [ctypes.c_void_p]),
    "mpeg1_decoder_get_write_ptr": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_uint]),
    "mpeg1_decoder_get_index": (ctypes.c_int, [ctypes.c_void_p]),
    "mpeg1_decoder_set_index": (None, [ctypes.c_void_p, ctypes.c_uint]),
    "mpeg1_decoder_did_write": (None, [ctypes.c_void_p, ctypes.c_uint]),
    "mpeg1_decoder_has_sequence_header": (ctypes.c_int, [ctypes.c_void_p]),
    "mpeg1_decoder_get_frame_rate": (ctypes.c_float, [ctypes.c_void_p]),
    "mpeg1_decoder_get_coded_size": (ctypes.c_int, [ctypes.c_void_p]),
    "mpeg1_decoder_get_width": (ctypes.c_int, [ctypes.c_void_p]),
    "mpeg1_decoder_get_height": (ctypes.c_int, [ctypes.c_void_p]),
    "mpeg1_decoder_get_y_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "mpeg1_decoder_get_cr_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "mpeg1_decoder_get_cb_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "mpeg1_decoder_decode": (ctypes.c_bool, [ctypes.c_void_p]),
}


def bind_mpeg1_abi(lib):
    for name, (res, args) in MPEG1_ABI.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library(path):
    """dlopen ``path`` with RTLD_LOCAL (all three libraries export the same symbol names)."""
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
    return bind_mpeg1_abi(ctypes.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW))


_product = None


def product_library():
    """The CUDA product library.  Fails loudly when the extension has not been built: there is
    no CPU fallback anywhere in the product path."""
    global _product
    if _product is None:
        _product = load_library(PRODUCT_LIB)
        from . import batch as _batch  # binds the batch-extension symbols on the same handle
        _batch.bind_batch_abi(_product)
    return _product


def bind_host_to_device(device):
    """Bind this thread (and the threads / pinned pages it creates from now on) to the CPUs of the NUMA
    node of CUDA device `device` (jsmpeg_b200_bind_host_to_device).  Returns {"node", "cpus"}."""
    node = ctypes.c_int(-1)
    n = product_library().jsmpeg_b200_bind_host_to_device(int(device), ctypes.byref(node))
    return {"node": node.value if n > 0 else None, "cpus": n if n > 0 else None, "bound": n > 0}
