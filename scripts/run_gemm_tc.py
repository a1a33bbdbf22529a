"""Launch the tensor-core GEMM of one layer a few times at bench size (for ncu captures)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import cuda_model
from coponerf_b200 import _lib
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # 1: A image, 2: out image
lib = _lib.load(); eng = cuda_model().engine()
K = [848, 832, 832, 832, 128, 128, 128][layer]
N = [832, 416, 416, 128, 128, 128, 128][layer]
A = torch.randn(M, K, device="cuda"); C = torch.empty(M, N, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    _lib.check(lib.cpn_gemm_tc(p(eng.weights), layer, p(A), K, p(C), N, M, 1, mode, 1, K // 32 if mode & 2 else 1, st), "gemm_tc"); ev[i + 1].record()
torch.cuda.synchronize()
kk = [835, 832, 832, 832, 128, 128, 128][layer]
for i in range(reps):
    ms = ev[i].elapsed_time(ev[i + 1]); print(f"layer {layer} mode {mode} M={M}: {ms:.3f} ms  {2*M*N*kk/ms/1e9:.1f} TFLOP/s (x3 MMA passes)")
