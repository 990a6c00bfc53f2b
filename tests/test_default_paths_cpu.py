"""The default-option fast paths of the momentum / scalar edge physics
(edge_physics.h, template DEF: alpha = 0, alpha_upw = 1, hoUpwind = 1) must
give the bits of the general paths: the specialisation only drops products with
exact 1 and exact 0.  Checked on the host build of the product header (the
CUDA build may contract FMAs differently in the two paths; the GPU parity
tests hold both to the oracle at 1e-12)."""
import ctypes as C

import parity_util as pu


def test_default_option_paths_are_bit_identical():
    L = pu.emu_lib()
    L.emu_default_path_mismatches.restype = C.c_int64
    L.emu_default_path_mismatches.argtypes = [C.c_int64, C.c_uint64]
    for seed in (1, 20261018):
        assert L.emu_default_path_mismatches(100000, seed) == 0
