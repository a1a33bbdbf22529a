"""Golden vectors for the closing stage of UFC.forward(), produced by the UNMODIFIED reference functions
(models/aggregation.py: correlation_token, interpolate4d, soft_argmax, unnormalise_and_convert_mapping_to_flow)
on seeded token features. Runs only in the build container.

    python tests/golden/make_goldens_ufc.py

The 64^4 volume `c` (67 MB) is not stored; the files keep the four flow fields, 8192 seeded samples of `c` and its
mean / mean-square, which pin it tightly enough.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_goldens import import_reference  # noqa: E402
from coponerf_b200 import synth  # noqa: E402

# (name, sizes, out, batch, seed, feature sharpness)
CASES = [("ufc_tail_small", (4, 8, 16), 16, 2, 11, 1.0),
         ("ufc_tail_256", (16, 32, 64), 64, 1, 12, 1.0)]


def main():
    import_reference()
    from models import aggregation as ag
    for name, sizes, out, batch, seed, sharp in CASES:
        src, trg = synth.ufc_tail_features(sizes, batch, seed)
        with torch.no_grad():
            corr = [ag.correlation_token(s, t, (n, n)) for s, t, n in zip(src, trg, sizes)]
            up = [ag.interpolate4d(x, (out, out, out, out)) for x in corr]
            c = sum(up) / len(up)
            gx, gy = ag.soft_argmax(c.permute(0, 1, 4, 5, 2, 3).flatten(1, 3))
            f_t2s = torch.cat((gx, gy), dim=1)
            flow = ag.unnormalise_and_convert_mapping_to_flow(f_t2s)
            gx, gy = ag.soft_argmax(c.flatten(1, 3))
            f_s2t = torch.cat((gx, gy), dim=1)
            flow_flip = ag.unnormalise_and_convert_mapping_to_flow(f_s2t)
        rng = np.random.default_rng(99)
        idx = rng.integers(0, c.numel(), size=8192)
        rec = dict(flow=flow.numpy(), flow_flip=flow_flip.numpy(), flow_t_to_s=f_t2s.numpy(), flow_s_to_t=f_s2t.numpy(),
                   c_idx=idx, c_val=c.reshape(-1)[idx].numpy(), c_mean=np.float64(c.double().mean()),
                   c_sq=np.float64((c.double() ** 2).mean()),
                   meta=np.array(list(sizes) + [out, batch, seed], dtype=np.int64))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, "c range", float(c.min()), float(c.max()), "flow range", float(flow.min()), float(flow.max()))


def conv4d_goldens():
    """Encoder4D blocks of the unmodified reference (models/conv4d.py) on the seeded volumes of synth.conv4d_case."""
    from models.conv4d import Encoder4D
    for name, (B, chans, k, stride, pad, n) in synth.CONV4D_CASES.items():
        x, layers, stride, pad = synth.conv4d_case(name)
        nl = len(chans) - 1
        m = Encoder4D(chans, ((k,) * 4,) * nl, ((stride,) * 4,) * nl, ((pad,) * 4,) * nl, (1,) * nl).eval()
        for blk, p in zip(m.conv4d, layers):
            c4, gn = blk[0], blk[1]
            with torch.no_grad():
                c4.query_conv.weight.copy_(p["wq"]); c4.query_conv.bias.copy_(p["bq"])
                c4.supp_conv.weight.copy_(p["ws"]); c4.supp_conv.bias.copy_(p["bs"])
                gn.weight.copy_(p["gamma"]); gn.bias.copy_(p["beta"])
        with torch.no_grad():
            y = m(x)
        rng = np.random.default_rng(77)
        idx = rng.integers(0, y.numel(), size=16384)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), idx=idx, val=y.reshape(-1)[idx].numpy(),
                            mean=np.float64(y.double().mean()), sq=np.float64((y.double() ** 2).mean()),
                            shape=np.array(y.shape, dtype=np.int64))
        print(name, tuple(y.shape), "mean", float(y.mean()), "max", float(y.max()))


def linatt_goldens():
    """LinearAttention of the unmodified reference (models/aggregation.py:84-117)."""
    from models.aggregation import LinearAttention
    att = LinearAttention()
    for name in synth.LINATT_CASES:
        q, k, v = synth.linatt_case(name)
        with torch.no_grad():
            y = att(q, k, v)
        rng = np.random.default_rng(78)
        idx = rng.integers(0, y.numel(), size=16384)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), idx=idx, val=y.reshape(-1)[idx].numpy(),
                            mean=np.float64(y.double().mean()), sq=np.float64((y.double() ** 2).mean()),
                            shape=np.array(y.shape, dtype=np.int64))
        print(name, tuple(y.shape), "max", float(y.abs().max()))


if __name__ == "__main__":
    main()
    conv4d_goldens()
    linatt_goldens()
