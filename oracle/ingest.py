"""TEST INFRASTRUCTURE — CPU restatement of the image normalisation of the reference's data
pipeline: cub/code/data/data.py:134,152 (same lines in pennaction/code/data/data.py)

    o.astype(np.float32) * 2.0 / 255.0 - 1.0

numpy keeps the arithmetic in fp32 (python scalars do not upcast a float32 array): one
multiply, one correctly rounded divide, one subtract.  Pinned against tests/golden/ingest.npz
(the reference's expression itself, evaluated by tests/golden/make_golden_ingest.py)."""
import numpy as np
import torch


def images_from_uint8(o):
    """uint8 array / tensor of any shape -> fp32 tensor in [-1, 1]."""
    a = o.numpy() if isinstance(o, torch.Tensor) else np.asarray(o)
    assert a.dtype == np.uint8, a.dtype
    x = a.astype(np.float32)
    x = x * np.float32(2.0)
    x = x / np.float32(255.0)
    x = x - np.float32(1.0)
    return torch.from_numpy(x)
