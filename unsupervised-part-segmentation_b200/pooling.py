"""Drop-in for the mask-weighted pooling helpers: deepfashion/code/foo.py:287-307,462-498,
574-578 and baselines/unsupervised-disentangling/ops.py:182-193."""
from . import ops
from .model import unpool_features  # noqa: F401  (foo.py:462 has the same helper)


def pool_features(feature_map, mask):
    """deepfashion/code/foo.py:287-307 — [bs,h,w,parts*f'],[bs,h,w,parts] -> reduce_mean over
    (h,w) of feature*mask -> [bs,parts,f']."""
    bs, h, w, n_features = feature_map.shape
    mshape = list(mask.shape)
    assert mshape[0] == bs and mshape[1] == h and mshape[2] == w, mshape
    n_parts = mshape[3]
    assert n_features % n_parts == 0, (n_features, n_parts)
    output = ops.part_pool(feature_map, mask, True, 1.0 / float(h * w))
    out_shape = list(output.shape)
    assert len(out_shape) == 3, out_shape
    assert out_shape[0] == bs and out_shape[1] == n_parts and out_shape[2] == n_features / n_parts, out_shape
    return output


def pool_unpool_block(feature_map, pool_mask, unpool_mask, reshape=False):
    """deepfashion/code/foo.py:574-578."""
    local_app_features = pool_features(feature_map, pool_mask)
    injected_mask = unpool_features(local_app_features, unpool_mask, reshape=reshape)
    return local_app_features, injected_mask


def get_features(features, part_map, slim):
    """baselines/unsupervised-disentangling/ops.py:182-193 — einsum('bijf,bijk->bkf') (slim)
    or einsum('bijkf,bijk->bkf')."""
    if slim:
        return ops.part_pool(features, part_map, False, 1.0)
    b, h, w, k, f = features.shape
    return ops.part_pool(features.reshape(b, h, w, k * f), part_map, True, 1.0)


def part_mean_pool(image, mask):
    """Pooling tail of the appearance encoder (cub/code/SB_model48i/model.py:50-52:
    reduce_mean over H,W) applied to mask_parts(image, mask): [B,H,W,C],[B,H,W,K] -> [B,K,C]."""
    return ops.part_pool(image, mask, False, 1.0 / float(image.shape[1] * image.shape[2]))
