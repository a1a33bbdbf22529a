"""Time the native cost aggregation (UFC.forward over CudaOps) for one 256x256 pair next to the CPU restatement."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coponerf_b200 import synth, ufc_native
from coponerf_b200.ufc_ops import CudaOps
sd = {k: v.cuda() for k, v in synth.ufc_state_dict(0).items()}
feat = [f.cuda() for f in synth.ufc_inputs(0)]
ops = CudaOps()
for _ in range(3):
    ufc_native.ufc_forward(sd, feat, 2, ops)
torch.cuda.synchronize()
times = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ufc_native.ufc_forward(sd, feat, 2, ops); b.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
# the same forward captured once in a CUDA graph and replayed (about 900 launches per pair)
g = torch.cuda.CUDAGraph()
static_feat = [f.clone() for f in feat]
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    ufc_native.ufc_forward(sd, static_feat, 2, ops)
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(g):
    graph_out = ufc_native.ufc_forward(sd, static_feat, 2, ops)
g.replay(); torch.cuda.synchronize()
eager = ufc_native.ufc_forward(sd, feat, 2, ops)
torch.cuda.synchronize()
same = all(torch.equal(a, b) for a, b in zip(graph_out[1], eager[1])) and torch.equal(graph_out[2], eager[2])
gt = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    gt.append(a.elapsed_time(b))
out = {"graph_ms_best": min(gt), "graph_equals_eager": same, "stage": "UFC.forward (aggregation.py:509-562), 1 pair 256x256, native operators", "gpu_ms_best": min(times),
       "gpu_ms_all": times}
if "--cpu" in sys.argv:
    from oracle.ufc_ops_torch import TorchOps
    torch.set_num_threads(os.cpu_count())
    sdc = synth.ufc_state_dict(0); fc = synth.ufc_inputs(0)
    ufc_native.ufc_forward(sdc, fc, 2, TorchOps())
    t0 = time.perf_counter(); ufc_native.ufc_forward(sdc, fc, 2, TorchOps()); out["cpu_restatement_ms"] = (time.perf_counter() - t0) * 1e3
    out["cores"] = os.cpu_count()
print(json.dumps(out))
