"""TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md 8f N4: the hourglass decoder's first
convolution applied to the injected part map, restated the way the reference computes it —
materialise [B,h,w,F+K], then a 3x3 SAME convolution plus bias:

    injected = tf.concat([tf.reduce_sum(unpool_features(feat, mask), 3), mask], 3)   cub/code/SB_model48i/model.py:482-484
    h = nn.conv2d(injected, config[0])                                               model.py:96 (hourglass_model, `dd` at :485)
      = tf.nn.conv2d(x, V, [1,1,1,1], "SAME") + tf.reshape(b, [1,1,1,num_filters])   cub/code/nn.py:661-663

V is TensorFlow's HWIO filter [3,3,F+K,Co], b [Co].  tf.nn.conv2d is a cross-correlation:
out[y,x,o] = sum_{i,j,c} in[y+i-1, x+j-1, c] * V[i,j,c,o], zero padding outside the image.
"""
import torch

from . import parts


def conv2d_same(x, V, b):
    """cub/code/nn.py:661-663 — NHWC input, HWIO filter, stride 1, SAME, + bias."""
    kh, kw = V.shape[0], V.shape[1]
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), V.permute(3, 2, 0, 1), padding=(kh // 2, kw // 2))
    return y.permute(0, 2, 3, 1) + b.reshape(1, 1, 1, -1)


def inject_conv2d(feature_vectors, mask, V, b):
    """model.py:482-485 + :96 — conv2d(concat(sum_k unpool_features(feat, mask), mask)).
    feature_vectors [B,K,F], mask [B,h,w,K], V [3,3,F+K,Co], b [Co] -> [B,h,w,Co]."""
    return conv2d_same(parts.inject(feature_vectors, mask), V, b)


def inject_conv_table(feature_vectors, V):
    """The per-sample filter table the CUDA path folds feat into (a pure re-association of the sum above):
    G[b,tap,k,o] = sum_f feat[b,k,f] * V[tap,f,o] + V[tap,F+k,o],  tap = 3*i + j."""
    B, K, F = feature_vectors.shape
    Co = V.shape[3]
    Vt = V.reshape(9, F + K, Co)
    return torch.einsum("bkf,tfo->btko", feature_vectors, Vt[:, :F]) + Vt[None, :, F:, :]
