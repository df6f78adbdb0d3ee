import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from synth import synth_input, synth_state_dict  # noqa: E402,F401


def load_fixture(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, arrays


def template_from_meta(meta):
    """State-dict-shaped template (zeros) with the reference's keys, shapes and dtypes."""
    out = {}
    for k, shape in meta["shapes"].items():
        dt = torch.int64 if "int64" in meta["dtypes"][k] else torch.float32
        out[k] = torch.zeros(shape, dtype=dt)
    return out


def digest(t):
    f = t.detach().double().flatten().cpu()
    return [float(f.norm())] + [float(x) for x in f[:4]]


def digest_close(a, b, rtol, atol=1e-7):
    """Compares [norm, first 4 values] digests."""
    norm_ok = abs(a[0] - b[0]) <= rtol * max(abs(b[0]), atol) + atol
    scale = max(abs(b[0]), atol)
    vals_ok = all(abs(x - y) <= rtol * scale + atol for x, y in zip(a[1:], b[1:]))
    return norm_ok and vals_ok
