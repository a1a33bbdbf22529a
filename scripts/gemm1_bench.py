"""Time the query_encode_latent GEMM (layer 0, operand image in, operand image out) at bench size in its variants.

    python scripts/gemm1_bench.py [rows] [reps] [modes...]     modes: extra CPN_TC_* bits (0 default, 16 cta pairs, 8 cluster)
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import cuda_model
from coponerf_b200 import _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
modes = [int(x) for x in sys.argv[3:]] or [0]      # 0 persistent, 256 one tile per CTA, 16 CTA pairs, 512 persistent CTA pairs
lib = _lib.load(); eng = cuda_model().engine()
tiles = M // 128
g = torch.Generator(device="cuda").manual_seed(0)
# a plausible image: fp16 values in [-2, 2) in the hi plane, small bytes in the correction planes
A = torch.zeros(tiles * 27 * 16384, dtype=torch.uint8, device="cuda")
Av = A.view(tiles * 27, 16384)
Av[:, :8192] = (torch.rand(tiles * 27, 4096, device="cuda", generator=g) * 4 - 2).half().view(torch.uint8)
Av[:, 8192:] = torch.randint(0, 64, (tiles * 27, 8192), dtype=torch.uint8, device="cuda", generator=g)
C = torch.empty(tiles * 26 * 16384, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode in modes:
    ms = []
    for i in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.cpn_gemm_tc(p(eng.weights), 0, p(A), 0, p(C), 0, M, 1, _lib.TC_A_IMAGE | _lib.TC_OUT_IMAGE | mode, 1, 26, st), "gemm_tc")
        b.record(); torch.cuda.synchronize()
        if i >= 2: ms.append(a.elapsed_time(b))
    best, med = min(ms), sorted(ms)[len(ms) // 2]
    print(f"gemm1 mode={mode} M={M}: median {med:.3f} ms best {best:.3f} ms -> {2*M*832*835/med/1e9:.1f} TFLOP/s algorithmic "
          f"checksum {int(C.view(-1)[::65537].sum())}", flush=True)
