"""Deterministic synthetic weights: the same numbers on every machine and torch version (numpy PCG64), so fixtures
only need to store inputs and outputs. Shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import numpy as np
import torch


def synth_state_dict(template: dict, seed: int) -> dict:
    """Fills a state-dict-shaped template with reproducible values: conv/linear weights ~ N(0, 1/fan_in) scaled like
    a trained net, BatchNorm weight in [0.75, 1.25], biases small, running_var in [0.5, 1.5], PReLU slopes in
    [0.1, 0.4]; integer buffers (num_batches_tracked) become 0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for k, v in template.items():
        shape = tuple(v.shape)
        if not v.is_floating_point():
            out[k] = torch.zeros(shape, dtype=v.dtype)
            continue
        n = int(np.prod(shape)) if shape else 1
        if k.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, n)
        elif k.endswith("running_mean"):
            a = rng.normal(0, 0.1, n)
        elif len(shape) == 1 and n == 1:           # PReLU slope
            a = rng.uniform(0.1, 0.4, n)
        elif len(shape) == 1 and k.endswith("weight"):   # BatchNorm gamma
            a = rng.uniform(0.75, 1.25, n)
        elif len(shape) == 1:                       # biases
            a = rng.normal(0, 0.05, n)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.normal(0, 1.0 / np.sqrt(fan_in), n)
        out[k] = torch.from_numpy(a.astype(np.float32)).reshape(shape)
    return out


def synth_input(shape, seed: int) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.uniform(0, 1, int(np.prod(shape))).astype(np.float32)).reshape(shape)
