"""where do the warps of k_conv_tc4 wait in the 3x3x3 convolution?  (sdb_tc_set_debug counters, cycles per tile)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from stardist_b200 import _lib as L
lib = L.require_cuda()
dbg = torch.zeros((148, 8), dtype=torch.int64, device='cuda')
names = ["mma:acc_empty", "mma:a_full", "mma:b_full", "mma:total", "epi:acc_full", "epi:total", "tma:a_empty", "-"]


def run(d, h, w, c0, c1, cout):
    cin = c0 + c1
    xs = torch.randn((2, d, h, w, cin), device='cuda').half()
    ws = (torch.randn((2, 27, cout, cin), device='cuda') * 0.05).half()
    b = torch.zeros(cout, device='cuda')
    out = torch.empty((2, d, h, w, cout), dtype=torch.float16, device='cuda')
    src0 = xs[..., :c0].contiguous() if c0 else None; src1 = xs[..., c0:].contiguous()
    args = (L.ptr(src0[0]) if c0 else L.ptr(None), L.ptr(src0[1]) if c0 else L.ptr(None), c0, L.ptr(src1[0]), L.ptr(src1[1]), c1,
            d, h, w, L.ptr(ws[0]), L.ptr(ws[1]), 1.0, L.ptr(b), cout, 1, 0)
    call = lambda: L.check(lib.sdb_conv3x3x3_tc(*args, L.ptr(out[0]), L.ptr(out[1]), L.stream_ptr()))
    for _ in range(2): call()
    L.check(lib.sdb_tc_set_debug(L.ptr(dbg)))
    dbg.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    L.check(lib.sdb_tc_set_debug(L.ptr(None)))
    v = dbg.double().mean(0).cpu().numpy()
    tiles = d * ((w + 127) // 128) * ((h + 1) // 2) / 148.0
    fl = 2.0 * d * h * w * cin * cout * 27
    print("(%d,%d,%d) %d->%d: %.1f us = %.0f TFLOP/s, %.1f tiles/CTA | " % (d, h, w, cin, cout, 1e3 * e0.elapsed_time(e1), fl / (1e-3 * e0.elapsed_time(e1)) / 1e12, tiles) +
          "  ".join("%s %.0f" % (n, x / tiles) for n, x in zip(names[:7], v[:7])) + "  (cycles per tile)", flush=True)


for cfg in [(64, 256, 256, 0, 32, 32), (64, 256, 256, 32, 32, 32), (32, 128, 128, 0, 64, 64), (32, 128, 128, 64, 64, 64), (64, 256, 256, 0, 32, 128)]:
    run(*cfg)
L.check(lib.sdb_tc_error_check(L.stream_ptr()))
