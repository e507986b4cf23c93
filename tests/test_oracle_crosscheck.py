"""CPU: the MLP / Adam part of the oracle (`oracle/model.py`, PARITY UNPINNED: TensorFlow/Keras cannot be installed here and
the reference ships no vectors) checked against INDEPENDENT third-party implementations of the same published
semantics that do exist in this image. This is not a pin on the reference's own TF run -- it shows that the oracle's
arithmetic is the textbook one, as implemented by somebody else:

  * Adam: scikit-learn's `AdamOptimizer` uses the same formulation as Keras' OptimizerV2 Adam (non-amsgrad) --
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), update = -lr_t * m / (sqrt(v) + eps), epsilon OUTSIDE the bias
    correction -- which is NOT torch.optim.Adam's formulation (eps added after dividing sqrt(v) by sqrt(1 - b2^t));
    the test also shows that the two differ by more than the tolerance, so it can tell them apart.
  * Dense + ReLU trunk: scikit-learn's MLPRegressor forward pass (`x @ coefs + intercepts`, ReLU on hidden layers)
    and torch.nn.functional.linear on the oracle's dense_0..dense_4 (core/model.py:366-372).
  * gradients: torch.autograd of the oracle's forward vs central finite differences in fp64.
"""
import numpy as np
import pytest
import torch

from oracle import model as om


def _toy_state(seed=0):
    rng = np.random.default_rng(seed)
    w = {"a/kernel": rng.standard_normal((7, 5)).astype(np.float32), "a/bias": rng.standard_normal(5).astype(np.float32)}
    grads = [{k: (rng.standard_normal(v.shape) * 10.0 ** rng.integers(-6, 1)).astype(np.float32) for k, v in w.items()}
             for _ in range(6)]
    return w, grads


def test_adam_matches_sklearn_formulation_over_several_steps():
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    w, grads = _toy_state()
    names = list(w)
    start = 1234          # the learning-rate schedule is evaluated at the Keras iteration counter
    ours = {k: v.copy() for k, v in w.items()}
    m = {k: np.zeros_like(v) for k, v in w.items()}
    v = {k: np.zeros_like(vv) for k, vv in w.items()}
    ref = [w[k].astype(np.float64) for k in names]
    opt = AdamOptimizer(ref, learning_rate_init=1.0, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    it = start
    for s, g in enumerate(grads):
        # Keras: bias correction uses t = iterations + 1 counted from 0; ExponentialDecay uses the same counter. sklearn's
        # optimizer counts its own t from 1, so the bias correction lines up when both start fresh and the schedule is fed in.
        opt.learning_rate_init = om.exponential_decay_lr(it - start)
        upd = opt._get_updates([g[k].astype(np.float64) for k in names])
        ref = [p + u for p, u in zip(ref, upd)]
        it_before = it - start
        om.adam_step(ours, g, m, v, it_before)
        it += 1
        for k, r in zip(names, ref):
            assert np.abs(ours[k] - r).max() <= 2e-7 * max(1.0, np.abs(r).max()), (s, k)
    # ... and the check can tell Keras' formulation from torch.optim.Adam's (epsilon inside the bias correction)
    tw = [torch.tensor(w[k], dtype=torch.float64, requires_grad=True) for k in names]
    topt = torch.optim.Adam(tw, lr=om.exponential_decay_lr(0), betas=(0.9, 0.999), eps=1e-7)
    for g in grads[:1]:
        for p, k in zip(tw, names):
            p.grad = torch.tensor(g[k], dtype=torch.float64)
        topt.step()
    w1 = {k: vv.copy() for k, vv in w.items()}
    om.adam_step(w1, grads[0], {k: np.zeros_like(x) for k, x in w.items()}, {k: np.zeros_like(x) for k, x in w.items()}, 0)
    diff = max(np.abs(w1[k] - p.detach().numpy()).max() for p, k in zip(tw, names))
    assert diff > 1e-6, "tiny gradients (|g| ~ eps) must separate the two epsilon conventions"


def test_dense_relu_trunk_matches_sklearn_and_torch_linear():
    from sklearn.neural_network import MLPRegressor
    w = om.init_weights(3, bias_scale=0.1)
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-1, 1, size=(257, 3)).astype(np.float32)
    enc = om.positional_encode(torch.from_numpy(xyz), 10)
    # the oracle's trunk up to the skip connection: dense_0..dense_4, ReLU each (core/model.py:366-368)
    wt = om.to_torch(w)
    h = enc
    for i in range(5):
        h = torch.relu(h @ wt[f"coarse/dense_{i}/kernel"] + wt[f"coarse/dense_{i}/bias"])
    ours = h.numpy()
    # scikit-learn: hidden layers with ReLU, identity output -> compare the last HIDDEN activation by making dense_4
    # the last hidden layer and a 256x256 identity the output layer
    net = MLPRegressor(hidden_layer_sizes=(256,) * 5, activation="relu")
    net.n_layers_ = 7
    net.out_activation_ = "identity"
    net.coefs_ = [w[f"coarse/dense_{i}/kernel"].astype(np.float64) for i in range(5)] + [np.eye(256)]
    net.intercepts_ = [w[f"coarse/dense_{i}/bias"].astype(np.float64) for i in range(5)] + [np.zeros(256)]
    sk = net._forward_pass_fast(enc.numpy().astype(np.float64), check_input=False)
    assert np.abs(ours - sk).max() <= 2e-5 * max(1.0, np.abs(sk).max())
    # torch.nn.functional.linear takes the kernel transposed ([out,in]): Keras stores [in,out] (core/model.py:366)
    g = enc
    for i in range(5):
        g = torch.relu(torch.nn.functional.linear(g, wt[f"coarse/dense_{i}/kernel"].T.contiguous(), wt[f"coarse/dense_{i}/bias"]))
    assert np.abs(ours - g.numpy()).max() <= 1e-5


def test_oracle_gradients_match_finite_differences_fp64():
    """d(loss)/d(parameter) of the oracle's train-step loss (core/model.py:148-168) by autograd vs central differences."""
    rng = np.random.default_rng(5)
    B = 6
    w = om.init_weights(7, bias_scale=0.05, sigma_gain=30.0)
    d = rng.standard_normal((B, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    ro = (rng.uniform(-0.2, 0.2, size=(B, 3))).astype(np.float32)
    rd = d.astype(np.float32)
    near = np.full((B, 1), 0.4, np.float32); far = np.full((B, 1), 1.3, np.float32)
    gt = rng.uniform(0, 1, size=(B, 3)).astype(np.float32)
    u_c = rng.random((B, 8), dtype=np.float32); u_f = rng.random((B, 16), dtype=np.float32)
    kw = dict(N_coarse=8, N_fine=16, u_coarse=u_c, u_fine=u_f, white_bg=True, dtype=torch.float64)
    info, g = om.loss_and_grads(w, ro, rd, near, far, gt, **kw)
    checked = 0
    for name, idx in (("fine/dense_3/kernel", (5, 17)), ("fine/rgb/bias", (1,)), ("coarse/dense_9/kernel", (260, 3)),
                      ("coarse/sigma/kernel", (40, 0)), ("fine/dense_0/bias", (9,))):
        h = 1e-4
        vals = []
        for sgn in (+1, -1):
            w2 = {k: v.astype(np.float64).copy() for k, v in w.items()}
            w2[name][idx] += sgn * h
            # the coarse network also moves the fine samples, but that path carries no gradient (stop_gradient,
            # utils/ray_utils.py:377): for a coarse parameter the tape's gradient is d(coarse loss) alone
            key = "coarse_loss" if name.startswith("coarse") else "loss"
            vals.append(om.loss_and_grads(w2, ro, rd, near, far, gt, **kw)[0][key])
        fd = (vals[0] - vals[1]) / (2 * h)
        an = float(g[name][idx])
        assert abs(fd - an) <= 2e-5 + 2e-3 * abs(fd), (name, fd, an)
        checked += 1
    assert checked == 5
