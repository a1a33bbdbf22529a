#!/usr/bin/env python
"""Rendered rays/s of the CoPoNeRF render path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--config 2|4]

Workloads (BASELINE.json `configs`):
  N = 1   configs[1]: one RealEstate10K-shape 256x256 stereo pair, all 65 536 target rays, S = 64 samples per line.
  N > 1   configs[2]: a batch of 8 such pairs sharded over the ranks (8 / N pairs per rank, no data-path collective), one
          NCCL gather of the final pixels to rank 0. The total work is fixed, so the line says "scaling": "strong"; the rate
          (rays/s) is directly comparable with the N = 1 line. The line also carries "strong": ONE pair with its rays
          sharded over the N ranks (coponerf_b200/dist.py: get_z on rank 0, one broadcast of its outputs, one all-gather of
          rgb), the curve that can bend (SURVEY.md section 8(e) case 2), with a bit-exact check against the 1-GPU render.
  --config 4   configs[3]: 512x512 pair, S = 128, 262 144 rays, cost aggregation at feature sizes 32 / 64 / 128 with a
          128^4 correlation volume (`--stage pair`: the encoder / pose head of get_z are hard-wired to 256x256 in the reference).
A "step" (default `--stage full`) is the whole drop-in call on one batch, images in, pixels out: get_z() (ResNet-34 encoder
+ conv_map in PyTorch/cuDNN, which BASELINE.json says stays; cost aggregation, pose features and pose head on the sm_100a
operators) followed by forward(val=True) over every ray with the estimated pose. `--stage pair` starts from the encoder's
feature pyramid (cost aggregation + render), `--stage render` times the render half alone (z, rel_pose, flow given).

The default `full` step treats the timed steps as a stream of pairs (coponerf_b200/pipeline.py, CoPoNeRF.render_pairs): get_z of
the next pair runs on a second stream while the current pair renders. Every timed region starts and ends with nothing in flight,
so it holds exactly `steps` x pairs get_z calls and renders. `"serial"` in the line is the same measurement with the two stages
back to back on one stream (`--no-pipeline` makes that the headline).

`value`  : inputs resident in HBM, timed on the device with CUDA events, L2 flushed between steps.
`e2e`    : the same through the drop-in forward() with HOST inputs: pinned host -> device copies of the images / poses / uv
           and the device -> host read of rgb (+ the reference's pixel_val.cpu()) are inside the timed region.
`--impl reference`: the reference path on the host cores, rank 0 only: the unmodified reference when a checkout is present
           (COPONERF_REFERENCE or baseline/_ref: models/CoPoNeRF.py driven through the three import shims), else the CPU
           oracle port (oracle/render_oracle.py + oracle/pair_oracle.py), on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 256
S = 64
N_RAYS = H * W
FLOP_PER_RAY = 667.9e6                    # reference formulation, S = 64
DOMINANT_MACS_PER_ROW = 835 * 832         # query_encode_latent, one encoder row
CPU_SAMPLE_RAYS = 1024
BATCH_PAIRS = 8                           # BASELINE config 3: batch-8 pairs over the GPUs of one box
UFC_SIZES = (16, 32, 64)                  # feature sizes of the cost aggregation at 256 x 256
UFC_MIN_BYTES = 135e6                     # weights + features in + c + features out per pair at 256 x 256
CONFIG = 2


def set_config(cfg):
    """BASELINE config 2 (default: 256 x 256, S = 64, 65 536 rays) or config 4 (512 x 512, S = 128, 262 144 rays,
    cost aggregation at feature sizes 32 / 64 / 128 with a 128^4 correlation volume)."""
    global H, W, S, N_RAYS, FLOP_PER_RAY, CPU_SAMPLE_RAYS, UFC_SIZES, UFC_MIN_BYTES, CONFIG
    CONFIG = cfg
    if cfg == 4:
        H = W = 512
        S = 128
        N_RAYS = H * W
        FLOP_PER_RAY = 1334.8e6           # SURVEY.md 8(d), S = 128
        CPU_SAMPLE_RAYS = 256
        UFC_SIZES = (32, 64, 128)
        UFC_MIN_BYTES = 1.2e9             # c alone is 1.07 GB


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 4], help="BASELINE config: 2 = 256x256 / S=64 (default), "
                    "4 = 512x512 / S=128 / 128^4 cost volume (implies --stage pair)")
    ap.add_argument("--chunk-rays", type=int, default=2048)
    ap.add_argument("--lanes", type=int, default=2, help="chunks in flight on internal streams")
    ap.add_argument("--simt", action="store_true", help="fp32 CUDA-core GEMMs only (cross-check path)")
    ap.add_argument("--no-fold", action="store_true", help="keep query_encode_latent_2 / latent_value / key_map as three GEMMs")
    ap.add_argument("--early-v", action="store_true", help="form V per sample (GEMM over all sample rows) instead of the late readout")
    ap.add_argument("--no-gfold", action="store_true", help="per-ray chain for the round-2 query bias + two readouts")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the one-pair ray-sharded measurement")
    ap.add_argument("--no-pipeline", action="store_true", help="run get_z and the render of every pair back to back on one stream "
                    "(default: get_z of the next pair overlaps the render of the current one, coponerf_b200/pipeline.py)")
    ap.add_argument("--eager-get-z", action="store_true", help="launch get_z()'s kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--stage", default="full", choices=["full", "pair", "render"],
                    help="full: get_z (encoder, cost aggregation, pose) + render per step; pair: cost aggregation "
                         "(UFC) + render from a given feature pyramid; render: render half only (z given)")
    a = ap.parse_args()
    set_config(a.config)
    if a.config == 4 and a.stage == "full":
        a.stage = "pair"
    return a


def workload_name(stage):
    head = f"{H}x{W} stereo pair, {N_RAYS} rays, S={S}"
    return {
        "full": head + ": get_z (ResNet-34 encoder, cost aggregation, pose features + pose head) + forward(val=True), "
                       "images in, pixels out",
        "pair": head + f": cost aggregation (UFC, feature sizes {'/'.join(map(str, UFC_SIZES))}) + render; encoder and pose "
                       "head outputs are inputs",
        "render": head + " (render half: forward with z given)",
    }[stage]


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_root():
    """A checkout of the unmodified reference, if one is reachable (never the case on a gpurun box unless the driver put one
    under baseline/_ref)."""
    for p in (os.environ.get("COPONERF_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if p and os.path.exists(os.path.join(p, "models", "CoPoNeRF.py")):
            return p
    return None


def cpu_pair_stage(stage, seed, timings=None):
    """Seconds the CPU restatement needs for the per-pair work of one image (reference formulation, incl. the
    Python positional-encoding loop of backbone.py:269-273), and its outputs (z, rel_pose, flow) for `full`."""
    from coponerf_b200 import synth
    if stage == "full":
        from oracle import pair_oracle
        sd = synth.full_state_dict(0)
        inp = synth.make_input(H, W, None, seed=seed, pose_set="frontal")
        pair_oracle.get_z(sd, inp, fast_pos=True)        # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        res = pair_oracle.get_z(sd, inp, fast_pos=False, timings=timings)
        return time.perf_counter() - t0, res
    if stage == "pair":
        from oracle import ufc_forward_oracle
        sdc, pyc = synth.ufc_state_dict(0, UFC_SIZES), synth.ufc_inputs(seed, 1, UFC_SIZES)
        if CONFIG != 4:
            ufc_forward_oracle.ufc_forward(sdc, pyc, 2)      # warm-up (skipped at 512 x 512: one pass takes tens of seconds)
        t0 = time.perf_counter()
        ufc_forward_oracle.ufc_forward(sdc, pyc, 2)
        return time.perf_counter() - t0, None
    return 0.0, None


def workload(seed):
    from coponerf_b200 import synth
    inp = synth.make_input(H, W, None, seed=seed, pose_set="frontal")
    z, rel_pose, flow = synth.make_features(H, W, seed=seed)
    return inp, z, rel_pose, flow


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "measured (bf16 dense, sustained)",
                d.get("hbm_gbs", 6650.0))
    return 1400.0, "fallback (B200_PROFILING.md sustained)", 6650.0


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = f"/tmp/cpn_clocks_{os.getpid()}.csv"
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                out["sm_max_mhz"] = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def strided_sample(inp):
    import torch
    idx = torch.arange(0, N_RAYS, N_RAYS // CPU_SAMPLE_RAYS)[:CPU_SAMPLE_RAYS]
    sub = {"context": inp["context"], "query": dict(inp["query"])}
    sub["query"]["uv"] = inp["query"]["uv"][:, :, idx].contiguous()
    sub["query"]["rgb"] = inp["query"]["rgb"][:, :, idx].contiguous()
    return sub, idx


def run_reference(args, rank):
    """The reference arm: the reference's own CPU path on a bounded sample of the workload, all host threads."""
    if rank != 0:
        return
    import torch
    from coponerf_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp, z, rel_pose, flow = workload(10)
    sub, _ = strided_sample(inp)
    ref_root = reference_root() if (args.stage == "full" and CONFIG == 2) else None
    if ref_root is not None:
        # the UNMODIFIED reference (models/CoPoNeRF.py) through the three import shims of SURVEY.md section 8(c)
        os.environ["COPONERF_REFERENCE"] = ref_root
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_goldens import import_reference
        model = import_reference().CoPoNeRF(n_view=2).eval()
        model.load_state_dict(synth.full_state_dict(0), strict=True)
        with torch.no_grad():
            model.get_z(inp)                                  # warm-up
            t0 = time.perf_counter()
            zr, pr, fr = model.get_z(inp)
            t_pair = time.perf_counter() - t0

            def step():
                return model(sub, z=zr, rel_pose=pr, flow=fr, val=True)
            for _ in range(args.warmup):
                step()
            times = []
            for _ in range(args.steps):
                t0 = time.perf_counter()
                step()
                times.append(time.perf_counter() - t0)
        kind = "reference"
        what = f"unmodified reference ({ref_root}: models/CoPoNeRF.py get_z + forward(val=True)), torch CPU fp32"
    else:
        from oracle import render_oracle
        sd = synth.render_state_dict(0)
        t_pair, res = cpu_pair_stage(args.stage, 10)
        if res is not None:       # render from the restatement's own features / estimated pose / flows
            z, rel_pose, flow = res

        def step():
            return render_oracle.render_forward(sd, sub, z, rel_pose, flow, H, W, S, True, chunk=512)
        for _ in range(args.warmup):
            step()
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            step()
            times.append(time.perf_counter() - t0)
        kind = "port"
        what = "CPU oracle port of the reference path (oracle/), torch CPU fp32"
    dt = sum(times) / len(times)
    best = min(times)
    # whole-image rate: the per-pair work once + all rays at the sampled per-ray cost
    v = N_RAYS / (t_pair + (N_RAYS / CPU_SAMPLE_RAYS) * dt)
    sample = (f"{what}: {CPU_SAMPLE_RAYS} evenly strided rays of the {N_RAYS}-ray image per step, mean of "
              f"{args.steps} steps (best {best * 1e3:.0f} ms), extrapolated to the image; per-pair stage ({args.stage}) timed "
              f"once ({t_pair * 1e3:.0f} ms)")
    line = {
        "impl": "reference", "metric": "rendered rays/sec at 256x256 stereo", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.stage), "stage": args.stage, "baseline_config": CONFIG, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "cpu_model": cpu_model(), "kind": kind,
                         "sample": sample, "best_of_steps_value": N_RAYS / (t_pair + (N_RAYS / CPU_SAMPLE_RAYS) * best)},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from coponerf_b200 import _lib, synth
    from coponerf_b200.dist import gather_rays, shard_range
    from coponerf_b200.model import CoPoNeRF

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    model = CoPoNeRF(n_view=2, npoints=S, chunk_rays=args.chunk_rays, lanes=args.lanes)
    full = args.stage == "full"
    if full:
        model.load_state_dict(synth.full_state_dict(0), strict=True)
    else:
        model.load_state_dict(synth.render_state_dict(0), strict=False)
    model = model.to(dev).eval()
    model.graph_get_z = not args.eager_get_z
    model.H, model.W = H, W
    eng = model.engine()
    eng.flags = (_lib.FLAG_SIMT_ONLY if args.simt else 0) | (_lib.FLAG_NO_FOLD if args.no_fold else 0) | \
        (_lib.FLAG_EARLY_V if args.early_v else 0) | (_lib.FLAG_NO_GFOLD if args.no_gfold else 0)
    lib = _lib.load()

    # BASELINE config 3 at N > 1: a batch of 8 pairs over the ranks; config 2 / 4 at N = 1: one pair
    ppr = max(1, BATCH_PAIRS // world) if world > 1 else 1
    total_pairs = ppr * world
    seeds = [10 + rank * ppr + i for i in range(ppr)]

    # ---- cost aggregation (per-pair stage): seeded UFC parameters and encoder pyramid
    with_ufc = args.stage == "pair"
    if with_ufc:
        from coponerf_b200 import ufc_native
        from coponerf_b200.ufc_ops import CudaOps
        ufc_ops = CudaOps()
        ufc_sd = {k: v.to(dev) for k, v in synth.ufc_state_dict(0, UFC_SIZES).items()}
        pyr_host = [[t.contiguous().pin_memory() for t in synth.ufc_inputs(sd_, 1, UFC_SIZES)] for sd_ in seeds]
        pyr_d = [[t.to(dev) for t in p] for p in pyr_host]

    # ---- this rank's pairs: host (pinned) and device copies
    pin = lambda t: t.contiguous().pin_memory()
    todev = lambda t: t.to(dev, non_blocking=True)
    host, z_host, flow_host, rel_host, inp_d, z_d, flow_d, rel_d, uv_d, cpu_case = [], [], [], [], [], [], [], [], [], []
    for sd_ in seeds:
        inp_h, z_h, rel_h, flow_h = workload(sd_)
        cpu_case.append((inp_h, z_h, rel_h, flow_h))
        host.append({"context": {k: pin(v) for k, v in inp_h["context"].items()},
                     "query": {k: pin(v) for k, v in inp_h["query"].items() if k != "rgb"}})
        z_host.append([pin(t) for t in z_h])
        flow_host.append(tuple(pin(t) for t in flow_h))
        rel_host.append(pin(rel_h))
        inp_d.append({"context": {k: todev(v) for k, v in host[-1]["context"].items()},
                      "query": {k: todev(v) for k, v in host[-1]["query"].items()}})
        z_d.append([todev(t) for t in z_host[-1]])
        flow_d.append(tuple(todev(t) for t in flow_host[-1]))
        rel_d.append(todev(rel_host[-1]))
        uv_d.append(inp_d[-1]["query"]["uv"].reshape(1, N_RAYS, 2))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    rgb_rank = torch.empty((ppr, 1, N_RAYS, 3), device=dev)            # this rank's pixels, gathered with ONE collective
    gather_buf = [torch.empty((ppr, 1, N_RAYS, 3), device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    state = {}
    ufc_events = []

    def render_pair(i, inp, host_side):
        """One pair: per-pair stage + render. Returns the output dict (device tensors)."""
        if full:       # images -> features, estimated pose, flows -> pixels
            if host_side:   # the call a user makes: forward(input, val=True) with z=None
                return model(inp, val=True)
            z, rel, flows = model.get_z(inp)
            state["z"], state["flow"], state["rel"] = z, flows, rel
            st = eng.prepare_pair(inp, z, rel, flows, H, W, True)
            return eng.render_rays(st, uv_d[i], S)
        if with_ufc:   # refined features + flows of this pair; conv_map (z[3]) comes from the encoder side
            pyr = [todev(t) for t in pyr_host[i]] if host_side else pyr_d[i]
            timing = state.get("time_ufc") and not host_side
            if timing:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            feats, flows, _c = ufc_native.ufc_forward(ufc_sd, pyr, 2, ufc_ops)
            if timing:
                e1.record()
                ufc_events.append((e0, e1))
            z = feats + [todev(z_host[i][3]) if host_side else z_d[i][3]]
            rel = todev(rel_host[i]) if host_side else rel_d[i]
            state["z"], state["flow"] = z, flows
            if host_side:
                return model(inp, z=z, rel_pose=rel, val=True, flow=flows)
            st = eng.prepare_pair(inp, z, rel, flows, H, W, True)
            return eng.render_rays(st, uv_d[i], S)
        if host_side:
            return model(inp, z=[todev(t) for t in z_host[i]], rel_pose=todev(rel_host[i]), val=True,
                         flow=tuple(todev(t) for t in flow_host[i]))
        st = eng.prepare_pair(inp, z_d[i], rel_d[i], flow_d[i], H, W, True)
        return eng.render_rays(st, uv_d[i], S)

    # ---- the default `full` step: a stream of pairs through the two-stage pipeline (coponerf_b200/pipeline.py). get_z of the
    # next pair (about 800 short kernels, one CUDA graph) runs on a second stream while the current pair renders. A timed
    # region starts with nothing in flight and ends with nothing in flight: `steps * ppr` get_z calls and renders inside it,
    # the first get_z exposed, the others overlapped.
    pipelined = full and not args.no_pipeline
    if pipelined:
        from coponerf_b200.pipeline import PairPipeline
        pipe = PairPipeline(model, priority_high=os.environ.get("CPN_PIPE_PRIORITY", "1") != "0")
    pstate = {"left": 0, "handle": None}

    def pipe_begin(n_steps):
        pstate["left"], pstate["handle"] = n_steps * ppr, None

    def pipe_pair(i, src):
        """get_z outputs of pair i (started by the previous call, or here at the start of a region); starts the next one."""
        if pstate["handle"] is None:
            pstate["handle"] = pipe.submit(src[i])
        got = pipe.take(pstate["handle"])
        pstate["left"] -= 1
        pstate["handle"] = pipe.submit(src[(i + 1) % ppr]) if pstate["left"] > 0 else None
        return got

    def device_step_pipelined():
        o = None
        for i in range(ppr):
            eng._feat_cache.clear()
            inp, z, rel, flows = pipe_pair(i, inp_d)
            state["z"], state["flow"], state["rel"] = z, flows, rel
            st = eng.prepare_pair(inp, z, rel, flows, H, W, True)
            o = eng.render_rays(st, uv_d[i], S)
            if world > 1:
                rgb_rank[i].copy_(o["rgb"][0])
        if world > 1:
            dist.gather(rgb_rank, gather_buf, dst=0)
        state["out"] = o
        return o

    def e2e_step_pipelined():      # host buffers in (copied on the get_z stream), the reference's host-side outputs out
        out = None
        for i in range(ppr):
            eng._feat_cache.clear()
            inp, z, rel, flows = pipe_pair(i, host)
            out = model(inp, z=z, rel_pose=rel, flow=flows, val=True)
            rgb_rank[i].copy_(out["rgb"][0])
        if world > 1:
            dist.gather(rgb_rank, gather_buf, dst=0)
        rgb_pinned.copy_(rgb_rank, non_blocking=True)
        return out

    def device_step():
        o = None
        for i in range(ppr):
            eng._feat_cache.clear()     # feature re-layout included every step (the cache would hide it): new pair state each image
            o = render_pair(i, inp_d[i], False)
            if world > 1:
                rgb_rank[i].copy_(o["rgb"][0])
        if world > 1:
            dist.gather(rgb_rank, gather_buf, dst=0)
        state["out"] = o
        return o

    rgb_pinned = torch.empty((ppr, 1, N_RAYS, 3)).pin_memory()

    def e2e_step():
        out = None
        for i in range(ppr):
            eng._feat_cache.clear()
            inp = {"context": {k: todev(v) for k, v in host[i]["context"].items()},
                   "query": {k: todev(v) for k, v in host[i]["query"].items()}}
            out = render_pair(i, inp, True)
            rgb_rank[i].copy_(out["rgb"][0])
        if world > 1:
            dist.gather(rgb_rank, gather_buf, dst=0)
        rgb_pinned.copy_(rgb_rank, non_blocking=True)
        return out

    nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)
    h2d = sum(nbytes(d.values()) for d in host[0].values())
    if not full:
        h2d += (nbytes(pyr_host[0]) + nbytes([z_host[0][3]])) if with_ufc else (nbytes(z_host[0]) + nbytes(flow_host[0]))
        h2d += rel_host[0].numel() * 4
    h2d *= ppr
    d2h = ppr * (N_RAYS * 3 * 4 + 2 * N_RAYS * S * 2 * 4)    # rgb + the reference's out['pixel_val'].cpu()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup, begin=None):
        if begin and warmup:
            begin(warmup)
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        barrier()
        torch.cuda.synchronize()
        if begin:
            begin(steps)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    clocks = ClockSampler(local_rank) if rank == 0 else None
    chunks = (N_RAYS + args.chunk_rays - 1) // args.chunk_rays
    # warm up first so the profiling session only sees the timed launches
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    ms_dev = timed(device_step, args.steps, 0)
    serial = None
    if pipelined:      # the headline is the pipelined stream of pairs; the back-to-back figure stays in the line beside it
        serial = {"value": total_pairs * N_RAYS * args.steps / (ms_dev * 1e-3), "ms_per_step": ms_dev / args.steps,
                  "what": "get_z then render of every pair back to back on one stream (--no-pipeline)"}
        ms_dev = timed(device_step_pipelined, args.steps, args.warmup, begin=pipe_begin)
    # roofline pass: the dominant kernel timed with CUDA events around each launch. Chunks are issued on one lane
    # here so that an event pair brackets that kernel alone (with lanes > 1 other chunks' kernels share the GPU
    # and the bracket would include their time); the step time of this pass gives the kernel's share.
    prof_steps = min(args.steps, 3)
    eng.lanes = 1
    device_step()
    torch.cuda.synchronize()
    _lib.check(lib.cpn_prof_begin(chunks * prof_steps * ppr), "cpn_prof_begin")
    state["time_ufc"] = with_ufc
    ms_prof = timed(device_step, prof_steps, 0)
    state["time_ufc"] = False
    dom_ms, dom_n = ctypes.c_float(0), ctypes.c_int(0)
    _lib.check(lib.cpn_prof_end(ctypes.byref(dom_ms), ctypes.byref(dom_n)), "cpn_prof_end")
    ufc_ms = [a.elapsed_time(b) for a, b in ufc_events]
    eng.lanes = args.lanes
    launches = args.steps * ppr * (eng.last_launch_count + 6)   # + 4 feature re-layouts, pair_setup, pair_prologue
    if with_ufc:
        launches += args.steps * ppr * ufc_ops.launches_per_forward
    if full:       # cost aggregation + pose operators of one get_z (the cuDNN encoder kernels are not ours: not counted)
        model.graph_get_z = False     # counted on one eager call; the timed steps replay the same kernels from a graph
        n0 = model._ufc_ops.launches
        model.get_z(inp_d[0])
        launches += args.steps * ppr * (model._ufc_ops.launches - n0)
        model.graph_get_z = not args.eager_get_z
    ms_e2e = timed(e2e_step, args.steps, 2)
    if pipelined:
        serial["e2e"] = {"value": total_pairs * N_RAYS * args.steps / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / args.steps}
        ms_e2e = timed(e2e_step_pipelined, args.steps, 2, begin=pipe_begin)

    # ---- N > 1: ONE pair, rays sharded over the ranks (SURVEY.md 8(e) case 2). Rank 0 runs get_z, one broadcast of a
    # flat buffer carries z / rel_pose / flows to the others, every rank renders its slice, one all-gather of rgb.
    strong = None
    if world > 1 and full and not args.no_strong:
        inp0_h, _, _, _ = workload(10)
        inp0 = {"context": {k: v.to(dev) for k, v in inp0_h["context"].items()},
                "query": {k: v.to(dev) for k, v in inp0_h["query"].items() if k != "rgb"}}
        uv0 = inp0["query"]["uv"].reshape(1, N_RAYS, 2)
        zt, relt, flt = model.get_z(inp0)                # shapes of the per-pair outputs (every rank, once)
        shapes = [tuple(t.shape) for t in zt] + [tuple(relt.shape)] + [tuple(t.shape) for t in flt]
        sizes = [math.prod(s) for s in shapes]
        flat = torch.empty(sum(sizes), device=dev)
        lo, hi = shard_range(N_RAYS, world, rank)

        def strong_step():
            eng._feat_cache.clear()
            if rank == 0:
                z, rel, flows = model.get_z(inp0)
                torch.cat([t.reshape(-1) for t in list(z) + [rel] + list(flows)], out=flat)
            dist.broadcast(flat, src=0)
            parts = [p.view(s) for p, s in zip(flat.split(sizes), shapes)]
            z, rel, flows = parts[:4], parts[4], tuple(parts[5:])
            st = eng.prepare_pair(inp0, z, rel, flows, H, W, True)
            o = eng.render_rays(st, uv0[:, lo:hi], S)
            state["strong_rgb"] = gather_rays(o["rgb"], N_RAYS, -2)
            return o

        ms_strong = timed(strong_step, args.steps, 2)

        def strong_step_redundant():     # every rank runs get_z itself (no broadcast): what SURVEY.md 8(e) proposes
            eng._feat_cache.clear()
            z, rel, flows = model.get_z(inp0)
            st = eng.prepare_pair(inp0, z, rel, flows, H, W, True)
            o = eng.render_rays(st, uv0[:, lo:hi], S)
            state["strong_rgb_r"] = gather_rays(o["rgb"], N_RAYS, -2)
            return o
        ms_strong_r = timed(strong_step_redundant, args.steps, 2)
        torch.cuda.synchronize()
        # parity: the gathered image against a single-GPU render of the same pair on this rank, bit for bit
        z, rel, flows = model.get_z(inp0)
        st = eng.prepare_pair(inp0, z, rel, flows, H, W, True)
        single = eng.render_rays(st, uv0, S)["rgb"]
        same = torch.tensor([int(torch.equal(single, state["strong_rgb_r"]))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        # the broadcast variant renders from rank 0's get_z: equal to rank 0's single-GPU render
        same_b = torch.tensor([int(torch.equal(single, state["strong_rgb"])) if rank == 0 else 1], device=dev)
        dist.all_reduce(same_b, op=dist.ReduceOp.MIN)
        strong = {"workload": "ONE 256x256 pair, 65536 rays sharded over the ranks: get_z on rank 0 + one broadcast of z / "
                              "rel_pose / flow (%.1f MB), each rank renders its contiguous ray slice, one all-gather of rgb"
                              % (flat.numel() * 4 / 1e6),
                  "value": N_RAYS * args.steps / (ms_strong * 1e-3), "unit": "rays/s", "ms_per_step": ms_strong / args.steps,
                  "redundant_get_z": {"value": N_RAYS * args.steps / (ms_strong_r * 1e-3), "ms_per_step": ms_strong_r / args.steps,
                                      "note": "every rank runs get_z itself, no broadcast"},
                  "bit_exact_vs_single_gpu_render": bool(same.item()) and bool(same_b.item()),
                  "what_is_left": "get_z (encoder, cost aggregation, pose) is not sharded and sits on the critical path of "
                                  "every step (Amdahl), plus tile quantisation of the last chunk per rank"}
    clk = clocks.stop() if clocks else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_rays = total_pairs * N_RAYS * args.steps
    value = total_rays / (ms_dev * 1e-3)
    e2e_value = total_rays / (ms_e2e * 1e-3)
    peak, peak_src, hbm_peak = peaks()
    rows_per_launch = 2 * 2 * S * min(args.chunk_rays, N_RAYS)       # encoder rows: 2 branches x 2 views x S per ray
    flop_per_launch = 2.0 * DOMINANT_MACS_PER_ROW * rows_per_launch
    achieved = flop_per_launch * dom_n.value / (dom_ms.value * 1e-3) / 1e12 if dom_ms.value > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath) and S == 64 and args.chunk_rays == 2048:
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("simt" if args.simt else "tc"), tj.get("source")
    line = {
        "metric": "rendered rays/sec at 256x256 stereo", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.stage), "baseline_config": CONFIG if world == 1 else 3, "stage": args.stage,
                   "get_z": "eager launches" if args.eager_get_z else "CUDA graph replay",
                   "pairs_per_step": total_pairs, "pairs_per_gpu": ppr, "chunk_rays": args.chunk_rays, "lanes": args.lanes,
                   "l2": "flushed between timed steps (256 MB write)",
                   "pipeline": ("get_z of pair k + 1 on a second stream while pair k renders (coponerf_b200/pipeline.py); every "
                                "timed region starts and ends with nothing in flight" if pipelined else "none: get_z and render back to back"),
                   "gemm_path": "simt-fp32" if args.simt else "tcgen05: fp16 head + two fp8 correction MMAs per product (e5m2 "
                                "activation planes, e4m3 weight planes, fp32 accumulate), persistent kernel"
                                + ("" if args.no_fold else "; query_encode_latent_2 folded into latent_value / key_map")
                                + ("" if (args.no_fold or args.early_v or args.simt) else "; attention reads out the hidden layer, latent_value per ray")
                                + ("" if (args.no_fold or args.early_v or args.simt or args.no_gfold) else "; round-2 query bias as a per-row linear map, one combined readout"),
                   "parallelism": (f"batch of {total_pairs} pairs sharded over {world} GPUs ({ppr} per rank), one NCCL gather of rgb"
                                   if world > 1 else "1 GPU")},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "query_encode_latent GEMM (835->832, ReLU)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": peak_src, "launches": dom_n.value,
                     "avg_launch_ms": dom_ms.value / max(dom_n.value, 1),
                     "share_of_step": dom_ms.value / ms_prof,
                     "measured": f"CUDA events around every launch in a {prof_steps}-step pass with lanes=1 "
                                 f"({ms_prof / prof_steps:.1f} ms/step)",
                     "flop_per_launch": flop_per_launch,
                     "whole_path_tflops": value * FLOP_PER_RAY / 1e12 / world,
                     # fp32 parity costs 2.0 tensor passes per product (fp16 head + two fp8 corrections at twice the fp16
                     # rate), so the reachable ceiling of this kernel is peak / 2
                     "tensor_passes_per_product": None if args.simt else 2.0,
                     "frac_of_parity_ceiling": None if args.simt else achieved / (peak / 2.0)},
        "clocks": clk,
    }
    if serial is not None:
        line["serial"] = serial
    if strong is not None:
        line["strong"] = strong
    if ufc_ms:      # cost aggregation timed on the device inside the roofline pass: an HBM-bound stage
        t_ufc = statistics.median(ufc_ms) * 1e-3
        line["roofline_ufc"] = {"bound": "hbm", "kernel": "cost aggregation (UFC.forward), all kernels of one pair",
                                "achieved": UFC_MIN_BYTES / t_ufc / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": UFC_MIN_BYTES / t_ufc / 1e9 / hbm_peak, "ms_per_pair": t_ufc * 1e3,
                                "algorithmic_bytes": UFC_MIN_BYTES, "traffic": None,
                                "launches_per_pair": ufc_ops.launches_per_forward}

    if world == 1 and not args.no_cpu_baseline:
        from oracle import render_oracle
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        inp_h, z_h, rel_h, flow_h = cpu_case[0]
        sub, idx = strided_sample(inp_h)
        sd = synth.render_state_dict(0)
        # the oracle renders from the same per-pair state the CUDA path used (the native UFC's outputs when the
        # cost aggregation is part of the step; its own parity is covered by tests/test_ufc_*_gpu.py)
        z_cpu = [t.detach().float().cpu().contiguous() for t in state["z"]] if (with_ufc or full) else z_h
        flow_cpu = tuple(t.detach().cpu() for t in state["flow"]) if (with_ufc or full) else flow_h
        rel_cpu = state["rel"].detach().cpu() if full else rel_h
        run = lambda: render_oracle.render_forward(sd, sub, z_cpu, rel_cpu, flow_cpu, H, W, S, True, chunk=512)
        ref = run()
        ts = []
        for _ in range(3 if CONFIG == 2 else 1):
            t0 = time.perf_counter()
            ref = run()
            ts.append(time.perf_counter() - t0)
        dt = min(ts)
        tm = {}
        t_pair, pair_res = cpu_pair_stage(args.stage, 10, tm)
        sample = (f"{CPU_SAMPLE_RAYS} evenly strided rays of the same image (chunks of 512), best of {len(ts)}, extrapolated to "
                  f"the image; per-pair stage ({args.stage}) timed once ({t_pair * 1e3:.0f} ms"
                  + (f": encoder {tm['encoder'] * 1e3:.0f}, cost aggregation {tm['ufc'] * 1e3:.0f}, pose {tm['pose'] * 1e3:.0f}"
                     if tm else "") + ")")
        line["cpu_baseline"] = {"value": N_RAYS / (t_pair + (N_RAYS / CPU_SAMPLE_RAYS) * dt), "unit": "rays/s",
                                "cores": cores, "cpu_model": cpu_model(), "kind": "port", "sample": sample}
        if pair_res is not None:     # get_z parity of this very step against the CPU restatement
            zc, pc, fc = pair_res
            line["pair_parity"] = {
                "rel_pose_max_abs_err": float((state["rel"].cpu() - pc).abs().max()),
                "z_max_rel_err": max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(state["z"], zc)),
                "flow_max_abs_err_px": max(float((a.cpu() - b).abs().max()) for a, b in zip(state["flow"][:2], fc[:2]))}
        got = state["out"]["rgb"][0, 0].cpu()[idx]
        want = ref["rgb"][0, 0]
        per_ray_err = (got - want).abs().max(dim=-1).values / want.abs().max()
        err = float(per_ray_err.max())
        # conditioning of the reference's own arithmetic on these rays: how far the oracle's rgb moves when the camera poses
        # move by about one fp32 ulp (the measure tests/cases.py gates with; two seeded perturbations here)
        import copy
        gsens = torch.Generator().manual_seed(1234)
        sens = torch.zeros(len(per_ray_err))
        for _ in range(2):
            pin = copy.deepcopy(sub)
            for grp in ("context", "query"):
                t = pin[grp]["cam2world"]
                pin[grp]["cam2world"] = t * (1 + 6e-8 * torch.randn(t.shape, generator=gsens))
            rp = rel_cpu * (1 + 6e-8 * torch.randn(rel_cpu.shape, generator=gsens))
            alt = render_oracle.render_forward(sd, pin, z_cpu, rp, flow_cpu, H, W, S, True, chunk=512)
            sens = torch.maximum(sens, (alt["rgb"][0, 0] - want).abs().max(dim=-1).values / want.abs().max())
        well = sens <= 1e-6
        mse = float(((got.clamp(-1, 1) - want.clamp(-1, 1)) ** 2).mean())
        line["parity"] = {"rgb_max_rel_err_vs_oracle": err, "rgb_median_rel_err_vs_oracle": float(per_ray_err.median()),
                          "rgb_p99_rel_err_vs_oracle": float(per_ray_err.kthvalue(int(0.99 * len(per_ray_err))).values),
                          "note": "the max sits on rays where the reference itself is ill-conditioned (DESIGN.md section 2)",
                          "oracle_one_ulp_pose_sensitivity": {"median": float(sens.median()), "max": float(sens.max()),
                                                              "rays_at_most_1e-6": float(well.float().mean())},
                          "rgb_max_rel_err_on_rays_with_sensitivity_at_most_1e-6":
                              float(per_ray_err[well].max()) if bool(well.any()) else None,
                          "psnr_vs_oracle_db": (-10 * math.log10(mse)) if mse > 0 else None,   # test.py:90-91 (clamped)
                          "psnr_unclamped_db": -10 * math.log10(max(float(((got - want) ** 2).mean()), 1e-30) /
                                                                float(want.abs().max()) ** 2),
                          "rays_checked": CPU_SAMPLE_RAYS}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
