"""Deterministic stand-in for pytorch-fid's InceptionV3 weight file (which cannot be downloaded here): every tensor of the
state dict is filled by closed-form arithmetic (no RNG), scaled so that activations stay O(1) through the network.  Shared by
tests/golden/gen_golden_eval.py (reference side) and tests/test_evaluation.py (our side)."""
import numpy as np
import torch


def synthetic_state(shapes):
    """shapes: {name: torch.Size} of a torchvision InceptionV3 (1008 classes, no aux head)"""
    out = {}
    for idx, name in enumerate(sorted(shapes)):
        shp = tuple(shapes[name])
        n = int(np.prod(shp)) if shp else 1
        wave = np.sin(np.arange(n, dtype=np.float64) * 0.37 + idx * 1.3)
        if name.endswith('num_batches_tracked'):
            out[name] = torch.zeros(shp, dtype=torch.long)
            continue
        if name.endswith('running_var'):
            v = 1.0 + 0.2 * wave ** 2
        elif name.endswith('running_mean'):
            v = 0.05 * wave
        elif name.endswith('bn.weight'):
            v = 1.0 + 0.1 * wave
        elif name.endswith('bias'):
            v = 0.05 * wave
        else:                                                   # conv / fc weights: He-like scale
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else n
            v = wave * np.sqrt(2.0 / fan_in) * 1.4
        out[name] = torch.from_numpy(v.reshape(shp).astype(np.float32))
    return out
