"""Generate tests/golden/inject_conv.npz by EXECUTING the reference's own functions (read from
/root/reference at generation time by `ast`, never copied) under tf1_shim:

  cub/code/SB_model48i/model.py: unpool_features :225 and the three statements at :482-484
      (reduce_sum over the part axis, concat with the decoding mask), spelled with the same tf calls;
  cub/code/nn.py: get_name :40, _conv2d :617 (3x3, stride 1, SAME, + bias) — the first layer of
      hourglass_model (model.py:96) that consumes the injected map (`dd`, model.py:485).

`tf.get_variable` hands `_conv2d` the V / b tensors of this fixture instead of creating variables.
Run in the build container only:  python tests/golden/make_golden_inject_conv.py
"""
import contextlib
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf1_shim as tf  # noqa: E402

REF = "/root/reference"


class TFVars(types.ModuleType):
    """tf1_shim plus the variable plumbing `_conv2d` touches."""

    def __init__(self):
        super().__init__("tf")
        self.__dict__.update(vars(tf))
        self.store = {}

    def variable_scope(self, name, **kw):
        return contextlib.nullcontext()

    def random_uniform_initializer(self, minval=0.0, maxval=1.0):
        return ("uniform", minval, maxval)

    def random_normal_initializer(self, stddev=1.0):
        return ("normal", stddev)

    def get_variable(self, name, shape=None, initializer=None, dtype=None, **kw):
        v = self.store[name]
        assert list(v.shape) == [int(s) for s in shape], (name, v.shape, shape)
        return v


def main():
    tfv = TFVars()
    ns = dict(tf=tfv, np=np, math=math, PARTS_DIM=3, FEATURE_DIM=4)
    tf.load_functions(f"{REF}/cub/code/nn.py", ["get_name", "_conv2d"], ns)
    tf.load_functions(f"{REF}/cub/code/SB_model48i/model.py", ["unpool_features"], ns)
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, B, H, W, K, F, Co, hard in (("a", 2, 8, 8, 4, 8, 16, True), ("b", 1, 6, 10, 16, 64, 32, True),
                                          ("c", 2, 5, 7, 5, 4, 8, False)):
        logits = torch.randn(B, H, W, K, generator=g)
        p = torch.softmax(logits, -1)
        if hard:
            # decoding mask = straight_through(hard_max(p), p): fl(fl(h - p) + p)   (nn.py:134-168)
            h = (p == p.max(-1, keepdim=True).values).float()
            mask = (h - p) + p
        else:
            mask = p                                  # a soft mask: the general (dense) path
        if tag == "a":
            mask[0, 3, 4] = 0.0; mask[0, 3, 4, 1] = 1.0; mask[0, 3, 4, 2] = 1.0   # an exact tie: two ones
            mask[1, 0, 0] = 0.0                                                     # an empty pixel
        mask = tf._t(mask).requires_grad_(True)
        feat = tf._t(torch.randn(B, K, F, generator=g)).requires_grad_(True)
        stdv = math.sqrt(1.0 / ((F + K) * 9))
        V = tf._t((torch.rand(3, 3, F + K, Co, generator=g) * 2 - 1) * stdv).requires_grad_(True)
        b = tf._t((torch.rand(Co, generator=g) * 2 - 1) * stdv).requires_grad_(True)
        tfv.store = {"V": V, "b": b}
        injected = ns["unpool_features"](feat, mask)
        injected = tfv.reduce_sum(injected, 3)
        injected = tfv.concat([injected, mask], axis=3)
        y = ns["_conv2d"](injected, Co)
        gy = torch.randn(y.shape, generator=g)
        dmask, dfeat, dV, db = torch.autograd.grad(y, [mask, feat, V, b], gy)
        out.update({f"{tag}_mask": mask, f"{tag}_feat": feat, f"{tag}_V": V, f"{tag}_b": b, f"{tag}_out": y,
                    f"{tag}_g_out": gy, f"{tag}_dmask": dmask, f"{tag}_dfeat": dfeat, f"{tag}_dV": dV, f"{tag}_db": db})
    np.savez_compressed(os.path.join(HERE, "inject_conv.npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print("wrote inject_conv.npz:", sorted(out)[:6], "...")


if __name__ == "__main__":
    main()
