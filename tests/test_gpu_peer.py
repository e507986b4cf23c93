"""GPU (one device): the peer-memory gradient exchange kernel (csrc/peer.cu) with a world of ONE rank -- the same kernel,
flags, epoch and fused Adam epilogue as in the multi-GPU step (tests/test_gpu_multi.py covers 2 ranks), runnable on
a single-GPU box. With one rank the sum over the ranks is the identity, so: the buffer must come back bit-identical,
the fused exchange + Adam must equal nerfb200_adam_step bit for bit, and both must replay from a CUDA graph."""
import ctypes as C

import numpy as np
import pytest
import torch

import nerf_tf2_b200 as nb
from nerf_tf2_b200 import _lib
from nerf_tf2_b200.model import _DeviceFloats

pytestmark = pytest.mark.gpu


def _peer(n, world=1, rank=0):
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.nerfb200_peer_create(world, rank, n, C.byref(h)), "peer_create")
    addr = C.c_void_p()
    _lib.check(lib.nerfb200_peer_buffer(h, C.byref(addr)), "peer_buffer")
    buf = torch.as_tensor(_DeviceFloats(addr.value, n), device="cuda")
    assert buf.data_ptr() == addr.value and buf.numel() == n and buf.dtype == torch.float32
    return lib, h, buf


def test_single_rank_exchange_is_the_identity_eager_and_graphed():
    n = 4 * 77777
    lib, h, buf = _peer(n)
    assert float(buf.abs().max()) == 0.0                           # the block arrives zeroed
    x = torch.randn(n, device="cuda")
    for it in range(3):
        buf.copy_(x * (it + 1))
        _lib.check(lib.nerfb200_peer_allreduce(h, _lib.stream_ptr()), "peer_allreduce")
        assert torch.equal(buf, x * (it + 1))
    g = torch.cuda.CUDAGraph()
    src = torch.empty_like(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        buf.copy_(src)
        _lib.check(lib.nerfb200_peer_allreduce(h, _lib.stream_ptr()), "peer_allreduce")
    for it in range(3):
        src.copy_(x * (it + 7))
        g.replay()
        assert torch.equal(buf, x * (it + 7))
    st = C.c_int(-1)
    _lib.check(lib.nerfb200_peer_status(h, C.byref(st)), "peer_status")
    assert st.value == 0
    del g, buf
    torch.cuda.synchronize()
    _lib.check(lib.nerfb200_peer_destroy(h), "peer_destroy")


@pytest.mark.parametrize("on_device_counter", [False, True])
def test_fused_exchange_adam_equals_adam_step_bit_for_bit(on_device_counter):
    n = _lib.PARAMS_TOTAL
    lib, h, buf = _peer(n + 4)
    gen = torch.Generator(device="cuda").manual_seed(3)
    grads = [torch.randn(n + 4, device="cuda", generator=gen) * 10.0 ** float(e) for e in (-1, -4, -7)]
    p0 = torch.randn(n, device="cuda", generator=gen)
    pa, ma, va = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb, mb, vb = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    start = 123456                                                  # deep enough into the schedule for lr to have decayed
    state_a = torch.tensor([start, 0], dtype=torch.int64, device="cuda") if on_device_counter else None
    state_b = torch.tensor([start, 0], dtype=torch.int64, device="cuda") if on_device_counter else None
    sp = lambda t: _lib.ptr(t, torch.int64, allow_none=True)
    for k, g in enumerate(grads):
        buf.copy_(g)
        _lib.check(lib.nerfb200_peer_allreduce_adam(h, n, _lib.ptr(pa), _lib.ptr(ma), _lib.ptr(va), start + k, sp(state_a),
                                                    _lib.stream_ptr()), "peer_allreduce_adam")
        _lib.check(lib.nerfb200_adam_step(n, _lib.ptr(pb), _lib.ptr(g), _lib.ptr(mb), _lib.ptr(vb), start + k, sp(state_b),
                                          _lib.stream_ptr()), "adam_step")
        if on_device_counter:
            _lib.check(lib.nerfb200_step_advance(sp(state_a), _lib.stream_ptr()), "step_advance")
            _lib.check(lib.nerfb200_step_advance(sp(state_b), _lib.stream_ptr()), "step_advance")
        assert torch.equal(buf, g)                                  # the exchanged buffer itself is untouched by Adam
        assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb), k
    assert not torch.equal(pa, p0)
    del buf
    torch.cuda.synchronize()
    _lib.check(lib.nerfb200_peer_destroy(h), "peer_destroy")


def test_exchange_refuses_to_launch_unconnected_or_misused():
    lib, h, buf = _peer(1024, world=2, rank=1)                      # two ranks, never connected
    assert lib.nerfb200_peer_allreduce(h, _lib.stream_ptr()) == 10003 and b"peer_connect" in lib.nerfb200_last_error()
    x = torch.zeros(2048, device="cuda")
    assert lib.nerfb200_peer_allreduce_adam(h, 2048, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), 0, None, _lib.stream_ptr()) == 10001
    assert lib.nerfb200_peer_allreduce_adam(h, 1022, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), 0, None, _lib.stream_ptr()) == 10001   # 4 | n
    assert lib.nerfb200_peer_allreduce_adam(h, 1024, C.c_void_p(x.data_ptr() + 4), _lib.ptr(x), _lib.ptr(x), 0, None,
                                            _lib.stream_ptr()) == 10001 and b"aligned" in lib.nerfb200_last_error()
    hd = C.create_string_buffer(64)
    _lib.check(lib.nerfb200_peer_handle(h, hd), "peer_handle")
    assert any(hd.raw)                                              # an IPC handle was produced
    blocks = (C.c_void_p * 2)(buf.data_ptr() - 4096, 0)
    a = C.c_void_p()
    assert lib.nerfb200_peer_attach(2, 0, 1024, blocks, None, C.byref(a)) == 10001      # a NULL block is refused
    del buf
    _lib.check(lib.nerfb200_peer_destroy(h), "peer_destroy")
    torch.cuda.synchronize()


def test_single_gpu_train_step_through_the_exchange_kernel_equals_plain_adam():
    """NeRF.train_step with the gradient buffer in a one-rank peer block and the fused exchange + Adam launch leaves the
    same parameters as the plain step (the data-parallel code path of model.py, on one GPU)."""
    H = W = 16
    sc = nb.scene.SyntheticScene(H, W, num_cameras=2)
    B = 256
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, H * W, (B,), generator=g, dtype=torch.int32).cuda()
    ro, rd = nb.ray_utils.get_rays_at(H, W, sc.K, sc.poses[0], ids)
    near = torch.full((B, 1), sc.near, device="cuda"); far = torch.full((B, 1), sc.far, device="cuda")
    rgb = torch.rand((B, 3), generator=g).cuda()
    batch = ((ro, rd, near, far), (rgb,))
    res = []
    for use_peer in (False, True):
        tn = nb.setup_model(nb.make_params({"system": {"white_bg": True}}, perturb=True), precision="bf16", train_precision="bf16",
                            seed=2, rng_seed=7)
        if use_peer:
            lib, h, buf = _peer(_lib.PARAMS_TOTAL + 4)
            tn._grad_buf, tn.flat_grads = buf, buf[:_lib.PARAMS_TOTAL]
            tn._peer, tn._peer_owner, tn.peer_mode, tn.world_size = h, buf, "ipc", 1
            tn._exchange_and_apply = lambda pending, step_state=None, tn=tn: tn.optimizer.exchange_and_apply(tn._peer, step_state=step_state)
        for _ in range(3):
            tn.train_step(batch)
        res.append((tn.flat_params.cpu().numpy(), tn.optimizer.m.cpu().numpy(), float(tn.last_loss.item())))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert abs(res[0][2] - res[1][2]) <= 1e-6 * abs(res[0][2])
