"""Stage 1b dequantises in sign-magnitude form (jsmpeg_b200/csrc/walk.cuh: dequant_sm) so that its loop needs
fewer ALU instructions; this pins it, exhaustively, to the reference's statements (src/mpeg1.js:794-807:
shift, +-1 for non-intra, multiply, arithmetic shift by 4, oddify toward zero, clamp) -- the device function
compiled for the host by the emulator build (tests/emu/walk_emu.cpp: emu_check_dequant)."""
import ctypes

import numpy as np

from test_walk_emu import emu_lib


def test_sign_magnitude_dequantisation_equals_the_reference_statements():
    lib = emu_lib()
    lib.emu_check_dequant.restype = ctypes.c_long
    lib.emu_check_dequant.argtypes = [ctypes.c_void_p]
    first = np.zeros(6, dtype=np.int32)
    bad = lib.emu_check_dequant(first.ctypes.data)
    assert bad == 0, f"{bad} mismatches; first: level {first[0]} qs {first[1]} Q {first[2]} intra {first[3]}: got {first[4]}, reference {first[5]}"
