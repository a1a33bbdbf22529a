"""Time the native closing stage of UFC (cpn_ufc_tail) at 256x256 shape next to the CPU oracle of the same stage."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coponerf_b200 import synth
from coponerf_b200.ufc import ufc_tail
from oracle import ufc_oracle
sizes, out = (16, 32, 64), 64
src, trg = synth.ufc_tail_features(sizes, 1, 12)
s_d, t_d = [t.cuda() for t in src], [t.cuda() for t in trg]
for _ in range(3):
    ufc_tail(s_d, t_d, sizes, out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    ufc_tail(s_d, t_d, sizes, out); ev[i + 1].record()
torch.cuda.synchronize()
gpu_ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
torch.set_num_threads(os.cpu_count())
ufc_oracle.ufc_tail(src, trg, sizes, out)
t0 = time.perf_counter(); ufc_oracle.ufc_tail(src, trg, sizes, out); cpu_ms = (time.perf_counter() - t0) * 1e3
# algorithmic bytes: features in (2 * 5376 * 256 * 4), c out (67.1 MB) + two soft-argmax reads of c
alg = 2 * 5376 * 256 * 4 + 3 * 4096 * 4096 * 4
print(json.dumps({"stage": "UFC closing stage (aggregation.py:527-561), 1 pair 256x256", "gpu_ms": gpu_ms, "cpu_oracle_ms": cpu_ms,
                  "cores": os.cpu_count(), "algorithmic_bytes": alg, "achieved_GBps": alg / gpu_ms / 1e6}))
