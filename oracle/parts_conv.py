"""TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md 8f N4, encoder side: the appearance encoder's first
convolution applied to the masked part images, restated the way the reference computes it — materialise
[B,h,w,K,3], fold it part-major, then a 3x3 SAME convolution plus bias on K*B images:

    view1_parts = mask_parts(self.augmented_views[1], self.encoding_mask)     cub/code/SB_model48i/model.py:478
    encode_parts(view1_parts, e_alpha) -> nn.apply_partwise(part_image, encoder)   model.py:214-222, cub/code/nn.py:81-113
    encoder_model: h = nn.conv2d(x, config[0])                                 model.py:40 (encoder1, :359-362)
      = tf.nn.conv2d(x, V, [1,1,1,1], "SAME") + b                              cub/code/nn.py:661-663

The value compared is the conv output in the part-major batch layout apply_partwise hands the encoder:
[K*B,h,w,Co], row k*B+b.
"""
from . import parts
from .inject_conv import conv2d_same


def fold_partmajor(part_image):
    """cub/code/nn.py:100-103 — [b,h,w,K,f] -> [K*b,h,w,f], row k*b_size+b."""
    b, h, w, k, f = part_image.shape
    return part_image.permute(3, 0, 1, 2, 4).reshape(k * b, h, w, f)


def parts_conv2d(image, mask, V, b):
    """image [B,h,w,3], mask [B,h,w,K], V [3,3,3,Co], b [Co] -> [K*B,h,w,Co]."""
    return conv2d_same(fold_partmajor(parts.mask_parts(image, mask)), V, b)
