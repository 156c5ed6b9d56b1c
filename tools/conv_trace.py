"""Per-role timeline of CTA 0 for one conv launch (debug tool)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from conv_bench import SHAPES  # noqa: E402
from hrnet_b200 import _lib  # noqa: E402
from hrnet_b200.ops import ConvLayer, PF8  # noqa: E402

name = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hw, cin, cout, k, stride, relu, use_res = SHAPES[name]
dev = torch.device("cuda")
w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
layer = ConvLayer(w, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), stride=stride, relu=relu)
x = PF8(N, cin, hw * stride, hw * stride); x.buf.normal_()
out = PF8(N, cout, hw, hw)
res = PF8(N, cout, hw, hw) if use_res else None
prm = layer.params(x, out, res)
lib = _lib.lib()
for _ in range(3):
    _lib.check(lib.hrnb_conv(C.byref(prm), _lib.stream_ptr()))
torch.cuda.synchronize()
buf = torch.zeros(5 * 64, dtype=torch.int64, device=dev)
lib.hrnb_debug_trace(C.c_void_p(buf.data_ptr()))
_lib.check(lib.hrnb_conv(C.byref(prm), _lib.stream_ptr()))
torch.cuda.synchronize()
lib.hrnb_debug_trace(None)
t = buf.cpu().view(5, 32, 2)
t0 = int(t[t > 0].min())
print("shape", name, "N", N, "BN", prm.BN, "MB", prm.MB, "KC", prm.KC, "(cycles relative to first event)")
print("iter | prod: start  issued | mma: top  got_tmem  got_AB  done_issue | epi: wait  got_acc  end")
for i in range(16):
    r = lambda a, b, c: (int(t[a, b, c]) - t0) if int(t[a, b, c]) else -1
    if not int(t[1, i, 0]):
        break
    print("%4d | %7d %7d | %7d %7d %7d %7d | %7d %7d %7d" % (i, r(0, i, 0), r(0, i, 1), r(1, i, 0), r(2, i, 0), r(2, i, 1),
                                                       r(1, i, 1), r(3, i, 0), r(3, i, 1), r(4, i, 0)))
