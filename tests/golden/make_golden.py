"""Generate tests/golden/*.npz by EXECUTING the reference's own functions (read from
/root/reference at generation time) under tf1_shim.  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; it only sees the committed .npz files.
Reference functions executed (file:line of the definitions):
  baselines/unsupervised-disentangling/transformations.py: tf_rotation_mat :5, tps_parameters :17,
      make_input_tps_param :59, ThinPlateSpline :93
  cub/code/nn.py: softmax :58, spatial_softmax :65, apply_partwise :81, hard_max_straight_through :118,
      hard_max :134, straight_through_estimator :154, probs_to_mu_sigma :1541, mask2hotmask :2086,
      unpool_features_gathered :2469
  cub/code/SB_model48i/model.py: categorical_kl :21, mask_parts :176, encode_parts :214, unpool_features :225
  deepfashion/code/foo.py: pool_features :287, unpool_features :462, pool_unpool_block :574
  baselines/unsupervised-disentangling/ops.py: get_features :182
"""
import os
import sys
import types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf1_shim as tf  # noqa: E402

REF = "/root/reference"


class DotMap(dict):
    __getattr__ = dict.__getitem__


def load_reference():
    ns_tps = dict(tf=tf, np=np, DotMap=DotMap)
    tf.load_functions(f"{REF}/baselines/unsupervised-disentangling/transformations.py",
                      ["tf_rotation_mat", "tps_parameters", "make_input_tps_param", "ThinPlateSpline"],
                      ns_tps)
    ns_nn = dict(tf=tf, np=np, deprecated=lambda **kw: (lambda f: f))
    tf.load_functions(f"{REF}/cub/code/nn.py",
                      ["softmax", "spatial_softmax", "apply_partwise", "hard_max_straight_through",
                       "hard_max", "straight_through_estimator", "mask2hotmask",
                       "unpool_features_gathered", "probs_to_mu_sigma"], ns_nn)
    nn = types.SimpleNamespace(**{k: v for k, v in ns_nn.items() if callable(v)})
    ns_model = dict(tf=tf, np=np, nn=nn, PARTS_DIM=3, FEATURE_DIM=4)
    tf.load_functions(f"{REF}/cub/code/SB_model48i/model.py",
                      ["mask_parts", "encode_parts", "unpool_features"], ns_model)
    ns_foo = dict(tf=tf, np=np, nn=nn)
    tf.load_functions(f"{REF}/deepfashion/code/foo.py",
                      ["pool_features", "unpool_features", "pool_unpool_block"], ns_foo)
    tf.load_functions(f"{REF}/cub/code/SB_model48i/model.py", ["categorical_kl"], ns_model)
    ns_ops = dict(tf=tf, np=np, wrappy=lambda f: f)
    tf.load_functions(f"{REF}/baselines/unsupervised-disentangling/ops.py", ["get_features"], ns_ops)
    return ns_tps, nn, ns_model, ns_foo, ns_ops


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().as_subclass(torch.Tensor).numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def T(x):
    return tf._t(x)


CUB_TPS = dict(scal=0.8, tps_scal=0.15, rot_scal=0.2, off_scal=0.2, scal_var=0.1, rescal=1.0)
PENN_TPS = dict(scal=0.95, tps_scal=0.08, rot_scal=0.05, off_scal=0.2, scal_var=0.05, rescal=1.0)


def gen_tps(ns, suffix=""):
    """suffix "_inv64": the same draws with the shim's matrix_inverse done in float64 (tf1_shim.MATRIX_INVERSE_FP64)."""
    g = torch.Generator().manual_seed(1234)
    tf.set_generator(g)
    tf.MATRIX_INVERSE_FP64 = suffix == "_inv64"
    for tag, kw, B, S, C in (("cub", CUB_TPS, 4, 16, 3), ("penn", PENN_TPS, 3, 24, 3),
                             ("big", dict(CUB_TPS, tps_scal=0.6, off_scal=0.6), 4, 16, 2)):
        prm = ns["tps_parameters"](B, **kw)
        coord, t_vector = ns["make_input_tps_param"](prm)
        U = T(torch.rand(B, S, S, C, generator=g) * 2 - 1).requires_grad_(True)
        tf.TAPS.clear()
        out, t_arr = ns["ThinPlateSpline"](U, coord, t_vector, S, C)
        W = tf.TAPS["matrix_inverse_in"][0]
        G = torch.randn(out.shape, generator=g)
        (dU,) = torch.autograd.grad(out, U, G)
        if suffix:   # inputs are those of tps_{tag}.npz (same generator sequence): only the results are stored
            npz(f"tps_{tag}{suffix}.npz", out=out, t_arr=t_arr, dU=dU)
            continue
        npz(f"tps_{tag}.npz", U=U, coord=coord, t_vector=t_vector, out=out, t_arr=t_arr, W=W,
            G=G, dU=dU, p_coord=prm.coord, p_vector=prm.vector, p_offset=prm.offset,
            p_offset_2=prm.offset_2, p_t_scal=prm.t_scal, p_rot_mat=prm.rot_mat)
    if suffix:
        tf.MATRIX_INVERSE_FP64 = False
        return
    # identity warp: unperturbed control points, zero displacement (SURVEY 8c vector 1)
    base = torch.tensor([[[-0.5, -0.5], [0.5, -0.5], [-0.5, 0.5], [0.5, 0.5],
                          [0.2, -0.2], [-0.2, 0.2], [0.2, 0.2], [-0.2, -0.2]]])
    coord = T(base.repeat(2, 1, 1))
    vec = T(torch.zeros(2, 8, 2))
    U = T(torch.rand(2, 16, 16, 3, generator=g) * 2 - 1)
    out, t_arr = ns["ThinPlateSpline"](U, coord, vec, 16, 3)
    npz("tps_identity.npz", U=U, coord=coord, t_vector=vec, out=out, t_arr=t_arr)
    # crop branch of make_input_tps_param + move/scal branch of ThinPlateSpline
    prm = ns["tps_parameters"](2, **CUB_TPS)
    mp = T(torch.rand(2, 1, 2, generator=g) * 0.2 - 0.1)
    sp = T(torch.rand(2, 2, generator=g) * 0.2 + 0.9)
    coord, t_vector = ns["make_input_tps_param"](prm, mp, sp)
    out, t_arr = ns["ThinPlateSpline"](U, coord, t_vector, 16, 3, move=mp, scal=sp)
    npz("tps_move.npz", U=U, coord=coord, t_vector=t_vector, out=out, t_arr=t_arr, move=mp, scal=sp,
        p_coord=prm.coord, p_vector=prm.vector, p_offset=prm.offset, p_offset_2=prm.offset_2,
        p_t_scal=prm.t_scal, p_rot_mat=prm.rot_mat)


def gen_parts(nn, ns_model, ns_foo, ns_ops):
    g = torch.Generator().manual_seed(4321)
    for tag, B, S, K, F in (("k4", 2, 8, 4, 6), ("k16", 2, 8, 16, 8), ("k25", 1, 6, 25, 5)):
        l0 = T(torch.randn(B, S, S, K, generator=g)).requires_grad_(True)
        l1 = T(torch.randn(B, S, S, K, generator=g)).requires_grad_(True)
        img = T(torch.rand(B, S, S, 3, generator=g) * 2 - 1).requires_grad_(True)
        feat = T(torch.randn(B, K, F, generator=g)).requires_grad_(True)
        Wlin = torch.randn(3, F, generator=g)

        def encoder(x):                      # stand-in for e_alpha's tail (model.py:50-52)
            return tf.reshape(tf.matmul(tf.reduce_mean(x, [1, 2]), Wlin), (-1, 1, 1, F))

        # the chain of cub/code/SB_model48i/model.py:426-485, call for call
        m0 = nn.softmax(l0, spatial=False)
        m1 = nn.softmax(l1, spatial=False)
        m0_hard = nn.straight_through_estimator(nn.hard_max(m0, 3), m0)
        m1_hard = nn.straight_through_estimator(nn.hard_max(m1, 3), m1)
        labels0 = tf.argmax(m0, axis=3)
        view1_parts = ns_model["mask_parts"](img, m1_hard)
        enc = ns_model["encode_parts"](view1_parts, encoder)
        u5 = ns_model["unpool_features"](feat, m0_hard)
        inj = tf.concat([tf.reduce_sum(u5, 3), m0_hard], axis=3)
        pooled = ns_ops["get_features"](img, m1_hard, True) / float(S * S)
        outs = [inj, view1_parts, pooled, m0, m1]
        cots = [torch.randn(o.shape, generator=g) for o in outs]
        grads = torch.autograd.grad(outs, [l0, l1, feat, img], cots)
        extra = {}
        if K == 4:
            fm = T(torch.randn(B, S, S, K * 3, generator=g))
            extra["pf_fmap"] = fm
            extra["pf_out"] = ns_foo["pool_features"](fm, m1)
            la, inj5 = ns_foo["pool_unpool_block"](fm, m1, m0, reshape=True)
            extra["pub_la"], extra["pub_inj"] = la, inj5
            extra["gf_dense"] = ns_ops["get_features"](fm, m1, True)
            extra["hotmask"] = nn.mask2hotmask(m0, K)
            extra["gathered"] = nn.unpool_features_gathered(feat, tf.cast(labels0, tf.int32))
            extra["spatial"] = nn.softmax(l0, spatial=True)
            extra["hmst"] = nn.hard_max_straight_through(m0, 3)
            extra["apw_identity"] = nn.apply_partwise(view1_parts, lambda x: x)
        npz(f"parts_{tag}.npz", l0=l0, l1=l1, img=img, feat=feat, Wlin=Wlin, m0=m0, m1=m1,
            m0_hard=m0_hard, m1_hard=m1_hard, labels0=labels0, view1_parts=view1_parts, enc=enc,
            u5=u5, inj=inj, pooled=pooled, g_inj=cots[0], g_parts=cots[1], g_pooled=cots[2],
            g_m0=cots[3], g_m1=cots[4], dl0=grads[0], dl1=grads[1], dfeat=grads[2], dimg=grads[3],
            **extra)
    # exact ties and extreme logits (SURVEY 8c vectors 4, 5)
    y = T(torch.tensor([[[[0.2, 0.5, 0.5, 0.1]]]]))
    lt = T(torch.tensor([[[[1.0, 3.0, 3.0, -2.0], [80.0, -80.0, 0.0, 79.0]],
                          [[0.0, 0.0, 0.0, 0.0], [-5.0, -5.0, 7.0, 7.0]]]]))
    pt = nn.softmax(lt)
    npz("parts_ties.npz", y=y, y_hard=nn.hard_max(y, 3), y_arg=tf.argmax(y, axis=3),
        st=nn.straight_through_estimator(T(torch.tensor([1.0])), T(torch.tensor([0.3]))),
        lt=lt, pt=pt, pt_hard=nn.hard_max(pt, 3), pt_arg=tf.argmax(pt, axis=3))


def gen_stats(ns_nn_funcs, ns_model):
    """probs_to_mu_sigma (cub/code/nn.py:1541) and categorical_kl (model.py:21)
    on spatial-softmax densities, as at cub/code/SB_model48i/model.py:437-440,659-661,683-689."""
    g = torch.Generator().manual_seed(777)
    out = {}
    for tag, B, H, W, K in (("a", 2, 8, 8, 4), ("b", 1, 6, 10, 25)):
        logits = T(torch.randn(B, H, W, K, generator=g)).requires_grad_(True)
        dens = tf.reshape(tf._t(torch.softmax(logits.reshape(B, H * W, K), dim=1)), (B, H, W, K))   # sums to 1 over HW
        sf = T(torch.rand(B, K, generator=g) * 0.5 + 0.75)
        mu, sigma = ns_nn_funcs["probs_to_mu_sigma"](dens, sf)
        p = tf._t(torch.softmax(logits, dim=-1))
        kl = ns_model["categorical_kl"](p)
        g_mu = torch.randn(mu.shape, generator=g)
        g_sigma = torch.randn(sigma.shape, generator=g)
        (d_dens,) = torch.autograd.grad([mu, sigma], [dens], [g_mu, g_sigma], retain_graph=True)
        (d_p,) = torch.autograd.grad(kl, p, torch.tensor(1.0))
        out.update({f"{tag}_dens": dens, f"{tag}_sf": sf, f"{tag}_mu": mu, f"{tag}_sigma": sigma, f"{tag}_p": p,
                    f"{tag}_kl": kl, f"{tag}_g_mu": g_mu, f"{tag}_g_sigma": g_sigma, f"{tag}_d_dens": d_dens,
                    f"{tag}_d_p": d_p})
    npz("stats.npz", **out)


def gen_priors():
    """mask priors and sampling on the part logits / probabilities.
    Reference functions executed: cub/code/nn.py fd_kernel :1357, tf_grad :1366, tf_squared_grad :1374,
    mumford_shah :1381, edge_set :1389, MeanFieldDistribution :1395 (sample :1421, kl_improper_gmrf :1444,
    kl_tv :1453), mask2hotmask :2086, mask2rgb :2067; call sites cub/code/SB_model48i/model.py:420-421 (sample),
    :667-681 (weak cross entropy), :744-769 (Mumford-Shah sums, area cost), :1071 (GMRF prior)."""
    colors_holder = [None]
    ns = dict(tf=tf, np=np, make_mask_colors=lambda n_parts: colors_holder[0])
    tf.load_functions(f"{REF}/cub/code/nn.py",
                      ["difference1d", "fd_kernel", "tf_grad", "tf_squared_grad", "mumford_shah", "edge_set",
                       "MeanFieldDistribution", "mask2hotmask", "mask2rgb", "softmax", "spatial_softmax", "hard_max",
                       "straight_through_estimator"], ns)
    g = torch.Generator().manual_seed(2024)
    tf.set_generator(g)
    out = {}
    for tag, B, H, W, K in (("a", 2, 8, 8, 4), ("b", 1, 6, 10, 16)):
        logits = T(torch.randn(B, H, W, K, generator=g)).requires_grad_(True)
        p = tf._t(torch.softmax(logits, dim=-1)).detach().requires_grad_(True)
        # --- Mumford-Shah on the probabilities (model.py:744-746); lambda chosen so that both branches occur
        alpha, lam = 1.0, 2.0e-3
        r, smooth, contour = ns["mumford_shah"](p, alpha, lam)
        assert 0.1 < float((smooth > 0).float().mean()) < 0.9
        sums = [tf.reduce_sum(v, axis=(1, 2)) for v in (r, smooth, contour, p)]         # [B,K] each
        cots = [torch.randn(v.shape, generator=g) for v in (r, smooth, contour)]
        gsums = [torch.randn(v.shape, generator=g) for v in sums]
        (d_p_elem,) = torch.autograd.grad([r, smooth, contour], [p], cots, retain_graph=True)
        (d_p_sums,) = torch.autograd.grad(sums, [p], gsums, retain_graph=True)
        edges = ns["edge_set"](p, alpha, lam)
        # --- MeanFieldDistribution on the logits
        dist = ns["MeanFieldDistribution"](logits, K)
        tf.set_generator(torch.Generator().manual_seed(99))
        sample = dist.sample(noise_level=0.7)
        eps = torch.randn(B, H, W, K, generator=torch.Generator().manual_seed(99))      # the same draw
        gmrf = dist.kl_improper_gmrf()
        tv = dist.kl_tv()
        kl0 = dist.kl()
        (d_gmrf,) = torch.autograd.grad(gmrf, logits, torch.tensor(1.0), retain_graph=True)
        (d_kl0,) = torch.autograd.grad(kl0, logits, torch.tensor(1.0), retain_graph=True)
        tf.set_generator(g)
        # --- weak cross entropy (model.py:667-681), both entropy_func settings
        p_labels = ns["softmax"](logits, spatial=False)
        labels = ns["straight_through_estimator"](ns["hard_max"](p_labels, 3), p_labels)
        ce = tf.reduce_mean(tf.nn.softmax_cross_entropy_with_logits_v2(labels, logits, dim=3))
        ent = tf.reduce_mean(tf.nn.softmax_cross_entropy_with_logits_v2(p_labels, logits, dim=3))
        (d_ce,) = torch.autograd.grad(ce, logits, torch.tensor(1.0), retain_graph=True)
        (d_ent,) = torch.autograd.grad(ent, logits, torch.tensor(1.0), retain_graph=True)
        # --- mask2rgb with an explicit colour table (the reference's make_mask_colors is matplotlib's inferno
        #     LUT, not available here and not on the path: the table is an input)
        colors = torch.rand(K, 3, generator=g)
        colors_holder[0] = colors.numpy().astype(np.float64)
        rgb_hot = ns["mask2rgb"](p.detach(), True)
        rgb_soft = ns["mask2rgb"](p.detach(), False)
        out.update({f"{tag}_logits": logits, f"{tag}_p": p, f"{tag}_alpha": np.float32(alpha), f"{tag}_lam": np.float32(lam),
                    f"{tag}_r": r, f"{tag}_smooth": smooth, f"{tag}_contour": contour, f"{tag}_edges": edges,
                    f"{tag}_sum_r": sums[0], f"{tag}_sum_smooth": sums[1], f"{tag}_sum_contour": sums[2], f"{tag}_sum_p": sums[3],
                    f"{tag}_g_r": cots[0], f"{tag}_g_smooth": cots[1], f"{tag}_g_contour": cots[2],
                    f"{tag}_g_sum_r": gsums[0], f"{tag}_g_sum_smooth": gsums[1], f"{tag}_g_sum_contour": gsums[2],
                    f"{tag}_g_sum_p": gsums[3], f"{tag}_d_p_elem": d_p_elem, f"{tag}_d_p_sums": d_p_sums,
                    f"{tag}_eps": eps, f"{tag}_sample": sample, f"{tag}_gmrf": gmrf, f"{tag}_tv": tv, f"{tag}_kl0": kl0,
                    f"{tag}_d_gmrf": d_gmrf, f"{tag}_d_kl0": d_kl0, f"{tag}_ce": ce, f"{tag}_ent": ent, f"{tag}_d_ce": d_ce,
                    f"{tag}_d_ent": d_ent, f"{tag}_colors": colors, f"{tag}_rgb_hot": rgb_hot, f"{tag}_rgb_soft": rgb_soft})
    npz("priors.npz", **out)


if __name__ == "__main__":
    ns_tps, nn, ns_model, ns_foo, ns_ops = load_reference()
    gen_tps(ns_tps)
    gen_tps(ns_tps, "_inv64")
    gen_parts(nn, ns_model, ns_foo, ns_ops)
    gen_stats(vars(nn), ns_model)
    gen_priors()
