"""Wide transform (tcgen05) vs the FMA transform at the C4 shape (ogbn-products: 2.45 M rows, 256 -> 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meta_gcn_b200 import ops, _lib
import ctypes

dev = torch.device("cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2449029
for hi, ho in ((256, 256), (100, 256), (64, 64)):
    x = torch.randn(N, hi, device=dev)
    w = torch.randn(hi, ho, device=dev) / hi ** 0.5
    b = torch.randn(ho, device=dev)

    def t(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    y = torch.empty(N, ho, device=dev)
    lib = _lib.load()
    fma = lambda: _lib.check(lib.mgcn_linear_ex(x.data_ptr(), None, N, hi, w.data_ptr(), ho, 1, ho, b.data_ptr(), None, 1,
                                                None, y.data_ptr(), torch.cuda.current_stream().cuda_stream))
    ms_tc = t(lambda: ops.linear_impl(x, w, False, b, None, 1))
    ms_fma = t(fma)
    ms_ref = t(lambda: torch.relu(torch.addmm(b, x, w)))
    flop = 2.0 * N * hi * ho
    byts = 4.0 * N * (hi + ho)
    print(f"[{N} x {hi}] x [{hi} x {ho}]: tcgen05 3xTF32 {ms_tc:.3f} ms ({flop / ms_tc / 1e9:.0f} TFLOP/s fp32-equivalent, "
          f"{3 * flop / ms_tc / 1e9:.0f} TF32 TFLOP/s issued, {byts / ms_tc / 1e6:.0f} GB/s) | FMA kernel {ms_fma:.3f} ms | "
          f"torch addmm+relu (cuBLAS fp32) {ms_ref:.3f} ms")

# weight gradient x^T g: tcgen05 (dispatch of mgcn_linear_wgrad_ex) at the C4 shapes
for hi, ho in ((256, 256), (100, 256)):
    x = torch.randn(N, hi, device=dev)
    g = torch.randn(N, ho, device=dev)
    for _ in range(2):
        ops.linear_wgrad_impl(x, g, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.linear_wgrad_impl(x, g, False, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"wgrad [{N} x {hi}]^T [{N} x {ho}]: {ms:.3f} ms ({2 * N * hi * ho / ms / 1e9:.0f} TFLOP/s fp32-equivalent, "
          f"{4 * N * (hi + ho) / ms / 1e6:.0f} GB/s)")
