"""where do the warps of k_conv_tc4 wait?  (sdb_tc_set_debug counters, cycles per CTA averaged over the grid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from stardist_b200 import _lib as L
lib = L.require_cuda()
dbg = torch.zeros((148, 8), dtype=torch.int64, device='cuda')
names = ["mma:acc_empty", "mma:a_full", "mma:b_full", "mma:total", "epi:acc_full", "epi:total", "tma:a_empty", "-"]
def run(h, w, c0, c1, cout):
    cin = c0 + c1
    xs = torch.randn((2, 1, h, w, cin), device='cuda').half()
    ws = torch.randn((2, 9, cout, cin), device='cuda').half()
    b = torch.zeros(cout, device='cuda')
    out = torch.empty((2, 1, h, w, cout), dtype=torch.float16, device='cuda')
    src0 = xs[..., :c0].contiguous() if c0 else None; src1 = xs[..., c0:].contiguous()
    args = (L.ptr(src0[0]) if c0 else L.ptr(None), L.ptr(src0[1]) if c0 else L.ptr(None), c0, L.ptr(src1[0]), L.ptr(src1[1]), c1,
            1, h, w, L.ptr(ws[0]), L.ptr(ws[1]), 1.0, L.ptr(b), cout, 1, 0)
    L.check(lib.sdb_tc_set_variant(4))
    for _ in range(2): L.check(lib.sdb_conv3x3_tc(*args, L.ptr(out[0]), L.ptr(out[1]), L.stream_ptr()))
    L.check(lib.sdb_tc_set_debug(L.ptr(dbg)))
    dbg.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); L.check(lib.sdb_conv3x3_tc(*args, L.ptr(out[0]), L.ptr(out[1]), L.stream_ptr())); e1.record(); torch.cuda.synchronize()
    L.check(lib.sdb_tc_set_debug(L.ptr(None))); L.check(lib.sdb_tc_set_variant(0))
    d = dbg.double().mean(0).cpu().numpy()
    tiles = ((w + 127) // 128) * ((h + 1) // 2) / 148.0
    print("(%d,%d) %d->%d: %.1f us, %.1f tiles/CTA | " % (h, w, cin, cout, 1e3 * e0.elapsed_time(e1), tiles) +
          "  ".join("%s %.0f" % (n, v / tiles) for n, v in zip(names[:7], d[:7])) + "  (cycles per tile)")
for cfg in [(1024, 1024, 0, 32, 32), (1024, 1024, 32, 32, 32), (1024, 1024, 0, 32, 128), (512, 512, 0, 64, 64)]:
    run(*cfg)
L.check(lib.sdb_tc_error_check(L.stream_ptr()))

# fused features + heads
h = w = 1024; cin = 32
xs = torch.randn((2, 1, h, w, cin), device='cuda').half()
ws = torch.randn((2, 9, 128, cin), device='cuda').half()
b = torch.zeros(128, device='cuda'); hw = torch.randn((128, 36), device='cuda') * 0.05; hb = torch.zeros(36, device='cuda')
prob = torch.empty((1, h, w), device='cuda'); dist = torch.empty((1, h, w, 32), device='cuda')
def fused():
    L.check(lib.sdb_conv3x3_heads_tc(L.ptr(None), L.ptr(None), 0, L.ptr(xs[0]), L.ptr(xs[1]), cin, 1, h, w, L.ptr(ws[0]), L.ptr(ws[1]), 1.0, L.ptr(b),
                                     1, L.ptr(hw), L.ptr(hb), 32, L.ptr(prob), L.ptr(dist), L.stream_ptr()))
for _ in range(2): fused()
L.check(lib.sdb_tc_set_debug(L.ptr(dbg))); dbg.zero_()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); fused(); e1.record(); torch.cuda.synchronize()
L.check(lib.sdb_tc_set_debug(L.ptr(None)))
d = dbg.double().mean(0).cpu().numpy(); tiles = 8 * 512 / 148.0
print("fused 32->128+heads: %.1f us | " % (1e3 * e0.elapsed_time(e1)) + "  ".join("%s %.0f" % (n, v / tiles) for n, v in zip(names[:7], d[:7])) + "  (cycles per tile)")
L.check(lib.sdb_tc_error_check(L.stream_ptr()))
