"""Generate tests/golden/parts_conv.npz by EXECUTING the reference's own functions (read from
/root/reference at generation time by `ast`, never copied) under tf1_shim:

  cub/code/SB_model48i/model.py: mask_parts :176
  cub/code/nn.py: apply_partwise :81, get_name :40, _conv2d :617 (3x3, stride 1, SAME, + bias) — the first layer of
      encoder_model (model.py:40) that `encode_parts(view1_parts, e_alpha)` (model.py:478-479) runs on every part image.

apply_partwise returns [b,h,w,K,Co]; the fixture stores it in the part-major layout [K*b,h,w,Co] the encoder sees.
Run in the build container only:  python tests/golden/make_golden_parts_conv.py
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf1_shim as tf  # noqa: E402
from make_golden_inject_conv import REF, TFVars  # noqa: E402


def main():
    tfv = TFVars()
    ns = dict(tf=tfv, np=np, math=math, PARTS_DIM=3, FEATURE_DIM=4)
    tf.load_functions(f"{REF}/cub/code/nn.py", ["get_name", "_conv2d", "apply_partwise"], ns)
    tf.load_functions(f"{REF}/cub/code/SB_model48i/model.py", ["mask_parts"], ns)
    g = torch.Generator().manual_seed(78)
    out = {}
    for tag, B, H, W, K, Co, hard in (("a", 2, 8, 8, 4, 16, True), ("b", 1, 6, 10, 16, 32, True),
                                      ("c", 2, 5, 7, 5, 8, False)):
        logits = torch.randn(B, H, W, K, generator=g)
        p = torch.softmax(logits, -1)
        if hard:
            h = (p == p.max(-1, keepdim=True).values).float()
            mask = (h - p) + p                       # straight_through(hard_max(p), p)   (nn.py:134-168)
        else:
            mask = p
        if tag == "a":
            mask[0, 3, 4] = 0.0; mask[0, 3, 4, 1] = 1.0; mask[0, 3, 4, 2] = 1.0   # an exact tie: two ones
            mask[1, 0, 0] = 0.0                                                     # an empty pixel
        mask = tf._t(mask).requires_grad_(True)
        image = tf._t(torch.rand(B, H, W, 3, generator=g) * 2 - 1).requires_grad_(True)
        stdv = math.sqrt(1.0 / (3 * 9))
        V = tf._t((torch.rand(3, 3, 3, Co, generator=g) * 2 - 1) * stdv).requires_grad_(True)
        b = tf._t((torch.rand(Co, generator=g) * 2 - 1) * stdv).requires_grad_(True)
        tfv.store = {"V": V, "b": b}
        part_image = ns["mask_parts"](image, mask)
        y5 = ns["apply_partwise"](part_image, lambda x: ns["_conv2d"](x, Co))     # [B,H,W,K,Co]
        y = y5.permute(3, 0, 1, 2, 4).reshape(K * B, H, W, Co)
        gy = torch.randn(y.shape, generator=g)
        dmask, dimage, dV, db = torch.autograd.grad(y, [mask, image, V, b], gy)
        out.update({f"{tag}_mask": mask, f"{tag}_image": image, f"{tag}_V": V, f"{tag}_b": b, f"{tag}_out": y,
                    f"{tag}_g_out": gy, f"{tag}_dmask": dmask, f"{tag}_dimage": dimage, f"{tag}_dV": dV,
                    f"{tag}_db": db})
    np.savez_compressed(os.path.join(HERE, "parts_conv.npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print("wrote parts_conv.npz:", sorted(out)[:6], "...")


if __name__ == "__main__":
    main()
