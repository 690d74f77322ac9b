"""TMA delivery rate for 64-byte vs 128-byte inner rows (halo boxes of the conv kernels), no MMA"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from stardist_b200 import _lib as L
lib = L.require_cuda()
for (h, w, c, box_c, rows, loads) in [(1024, 1024, 32, 32, 4, 1), (1024, 1024, 32, 32, 4, 2), (1024, 1024, 64, 64, 4, 1), (1024, 1024, 64, 32, 4, 2),
                                      (1024, 1024, 64, 64, 4, 2), (1024, 1024, 32, 32, 10, 1), (1024, 1024, 64, 64, 6, 1), (1024, 1024, 128, 64, 4, 2)]:
    x = torch.randn((1, h, w, c), device='cuda').half()
    ms = ctypes.c_float(0)
    L.check(lib.sdb_tma_probe(L.ptr(x), h, w, c, box_c, rows, loads, 20, ctypes.byref(ms), L.stream_ptr()))
    L.check(lib.sdb_tc_error_check(L.stream_ptr()))
    tiles = ((w + 127) // 128) * ((h + rows - 3) // (rows - 2))
    nbytes = tiles * loads * box_c * 2 * 130 * rows
    nrows = tiles * loads * 130 * rows
    print("h=%d w=%d c=%d box_c=%d rows=%d loads/tile=%d: %.1f us  %.2f TB/s  %.0f rows/us/SM (%.1f clk/row)" %
          (h, w, c, box_c, rows, loads, 1e3 * ms.value, nbytes / ms.value / 1e9, nrows / (1e3 * ms.value) / 148, 148 * 1965 * 1e3 * ms.value / nrows / 1e3))
