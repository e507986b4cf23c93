"""
Host-side mirror of the reference's `nerf/core/model.py`: the `NeRF` model with its Keras-shaped
surface (`fit` / `evaluate` / `predict`, `train_step` / `test_step` / `predict_step`, `call`,
`forward`, `compile`, `trainable_variables`, `optimizer.variables()`), `setup_model`,
`get_coarse_or_fine_model` and `PositionalEncoder`. Everything numeric runs in the sm_100a
kernels behind the C ABI; torch only owns device memory, streams and the process group.

Parameters are ONE flat fp32 device buffer (coarse model then fine model, variables in Keras
`model.trainable_variables` order = dense_0..dense_9, rgb, sigma; kernels [in,out] row-major): the 48 `trainable_variables` are views into it, the
gradient is one flat buffer (ONE exchange per data-parallel step, SURVEY.md 8e: the library's peer-memory kernel, or an NCCL all-reduce)
and Adam is one fused launch.
"""
import ctypes as C
import functools
import math

import numpy as np
import torch

from . import _lib, ops, ray_utils
from ._lib import BF16, COARSE, FINE, FP16, FP32, TF32, PARAMS_PER_MODEL, PARAMS_TOTAL, check, load, ptr, stream_ptr
from .data import RayDataset

LAYER_NAMES = [f"dense_{i}" for i in range(10)] + ["rgb", "sigma"]
LAYER_SHAPES = {
    "dense_0": (63, 256), "dense_1": (256, 256), "dense_2": (256, 256), "dense_3": (256, 256),
    "dense_4": (256, 256), "dense_5": (319, 256), "dense_6": (256, 256), "dense_7": (256, 256),
    "sigma": (256, 1), "dense_8": (256, 256), "dense_9": (283, 128), "rgb": (128, 3),
}


def variable_names(model_name):
    """`{model}/dense_i/kernel:0`-style names without the `:0` (core/model.py:367-386)."""
    out = []
    for ln in LAYER_NAMES:
        out += [f"{model_name}/{ln}/kernel", f"{model_name}/{ln}/bias"]
    return out


def glorot_uniform_params(seed):
    """Keras defaults for Dense: glorot_uniform kernels, zero biases; flat [PARAMS_TOTAL] fp32."""
    rng = np.random.default_rng(seed)
    parts = []
    for _ in ("coarse", "fine"):
        for ln in LAYER_NAMES:
            fi, fo = LAYER_SHAPES[ln]
            lim = math.sqrt(6.0 / (fi + fo))
            parts.append(rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32).reshape(-1))
            parts.append(np.zeros((fo,), dtype=np.float32))
    flat = np.concatenate(parts)
    assert flat.shape[0] == PARAMS_TOTAL
    return flat


def _on_device(fn):
    """Runs a NeRF method with the model's GPU as the current device: the C ABI launches on the current
    device's current stream and the context is bound to the device it was created on."""
    @functools.wraps(fn)
    def inner(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return inner


class PositionalEncoder:
    """PositionalEncoder (core/model.py:289-332). Standalone layer object for API parity; inside
    the fused MLP kernel the encoding is computed on chip and never written to HBM."""

    def __init__(self, L, name=None):
        self.L = L
        self.name = name

    def __call__(self, x):
        return ray_utils.positional_encode(x.contiguous(), self.L)

    call = __call__


class Variable:
    """A named view into the flat parameter buffer (stands in for tf.Variable)."""

    def __init__(self, name, view, owner=None):
        self.name = name
        self._view = view
        self._owner = owner          # the NeRF whose packed tensor-core images go stale when this changes
        self.shape = tuple(view.shape)

    def numpy(self):
        return self._view.detach().cpu().numpy().copy()

    def value(self):
        return self._view

    def assign(self, arr):
        self._view.copy_(torch.as_tensor(np.asarray(arr), dtype=torch.float32).reshape(self.shape))
        if self._owner is not None:
            self._owner._dirty = True


class SubModel:
    """The coarse or the fine 8x256 MLP (get_coarse_or_fine_model, core/model.py:334-394):
    callable on (xyz[R,3], dirs[R,3]) -> (rgb[R,3], sigma[R,1])."""

    def __init__(self, nerf, which, name):
        self._nerf, self.which, self.name = nerf, which, name

    @property
    def trainable_variables(self):
        return self._nerf._variables[self.which * 24:(self.which + 1) * 24]

    variables = weights = trainable_variables

    def get_weights(self):
        return [v.numpy() for v in self.trainable_variables]

    def set_weights(self, arrays):
        assert len(arrays) == 24
        for v, a in zip(self.trainable_variables, arrays):
            v.assign(a)
        self._nerf._dirty = True

    def save_weights(self, path):
        """Model.save_weights: Keras `.h5` layout through h5lite, or `.npz` (checkpoint.py)."""
        from . import checkpoint
        if path.endswith(".h5") or path.endswith(".hdf5"):
            checkpoint.save_weights_h5(path, self)
        else:
            checkpoint.save_weights(path, self.trainable_variables)

    def load_weights(self, path):
        from . import checkpoint
        checkpoint.load_weights(path, self)

    def __call__(self, inputs, precision=None):
        xyz, dirs = inputs
        R = xyz.shape[0]
        t0 = torch.zeros((R, 1), device=xyz.device, dtype=torch.float32)
        # a row is a 1-sample ray with o = xyz, t = 0: o + 0*d == o exactly
        with torch.cuda.device(self._nerf.device):
            rgb, sigma = self._nerf._mlp(self.which, xyz.contiguous(), dirs.contiguous(), t0, precision=precision)
        return rgb, sigma.reshape(R, 1)


class _Adam:
    """Keras Adam + ExponentialDecay(5e-4, 500000, 0.1) state (core/model.py:413-418)."""

    def __init__(self, nerf):
        self._nerf = nerf
        self.iterations = 0
        self.m = torch.zeros_like(nerf.flat_params)
        self.v = torch.zeros_like(nerf.flat_params)

    def learning_rate(self, step=None):
        step = self.iterations if step is None else step
        return 5e-4 * (0.1 ** (step / 500000.0))

    def variables(self):
        """[iter, m x48, v x48] -- the order CustomSaver/set_everything rely on (core/ops.py:146-149)."""
        out = [np.int64(self.iterations)]
        for buf in (self.m, self.v):
            for var in self._nerf._variables:
                out.append(buf[var._ofs:var._ofs + var._n].reshape(var.shape).cpu().numpy().copy())
        return out

    def set_weights(self, weights):
        self.iterations = int(weights[0])
        vs = self._nerf._variables
        for k, buf in enumerate((self.m, self.v)):
            for i, var in enumerate(vs):
                arr = torch.as_tensor(np.asarray(weights[1 + k * len(vs) + i]), dtype=torch.float32)
                buf[var._ofs:var._ofs + var._n].copy_(arr.reshape(-1))

    def apply_gradients(self, flat_grads, step_state=None):
        """One fused Adam launch over the flat parameter block. `step_state` (device int64[2]): the iteration count is
        read on the device (CUDA-graph replays); the caller then advances both the device and the host counter."""
        n = self._nerf.flat_params.numel()
        with torch.cuda.device(self._nerf.device):
            check(load().nerfb200_adam_step(n, ptr(self._nerf.flat_params), ptr(flat_grads), ptr(self.m), ptr(self.v),
                                            self.iterations, ptr(step_state, torch.int64, allow_none=True), stream_ptr()),
                  "adam_step")
        if step_state is None:
            self.iterations += 1
        self._nerf._dirty = True

    def exchange_and_apply(self, peer, step_state=None):
        """Data parallel: the gradient exchange over peer memory and the Adam step as ONE launch
        (nerfb200_peer_allreduce_adam) -- same arithmetic as an all-reduce followed by apply_gradients."""
        n = self._nerf.flat_params.numel()
        with torch.cuda.device(self._nerf.device):
            check(load().nerfb200_peer_allreduce_adam(peer, n, ptr(self._nerf.flat_params), ptr(self.m), ptr(self.v),
                                                      self.iterations, ptr(step_state, torch.int64, allow_none=True),
                                                      stream_ptr()), "peer_allreduce_adam")
        if step_state is None:
            self.iterations += 1
        self._nerf._dirty = True


class _DeviceFloats:
    """Exposes `n` floats of device memory owned by the C library to torch (CUDA array interface)."""

    def __init__(self, address, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(address), False), "version": 2}


class History:
    def __init__(self):
        self.history = {}
        self.epoch = []


class NeRF:
    """NeRF(Model) (core/model.py:18-287)."""

    def __init__(self, params, precision="bf16", train_precision=None, seed=0, device=None,
                 rng_seed=0, render_chunk=32768, precise_last=True, cuda_graph=False):
        """`precision`: MLP arithmetic of forward/predict -- "bf16" (default), "fp16", "tf32" (tcgen05 tensor cores,
        fp32 accumulate) or "fp32" (CUDA-core check path). `train_precision` defaults to `precision` (bf16 for a
        tf32 model: tf32 is a render precision). `precise_last`: the tensor-core forwards recompute sigma of every
        ray's last sample with split operands (NERFB200_OPT_PRECISE_LAST): True = render forwards (default),
        "train" = training forwards too, False = never."""
        _lib.require_cuda()
        load()
        self.params = params
        self.white_bg = params.system.white_bg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = _lib.PRECISIONS[precision] if isinstance(precision, str) else precision
        tp = train_precision if train_precision is not None else ("bf16" if self.precision == TF32 else self.precision)
        self.train_precision = _lib.PRECISIONS[tp] if isinstance(tp, str) else tp
        # False / True (render forwards) / "train" (training forwards too: NERFB200_OPT_PRECISE_LAST = 2)
        self.precise_last = 2 if precise_last == "train" else int(bool(precise_last))
        self.rng_seed = rng_seed
        self.render_chunk = render_chunk
        self.val_cache = []
        with torch.cuda.device(self.device):
            self.flat_params = torch.from_numpy(glorot_uniform_params(seed)).to(self.device)
            # ONE flat buffer carries everything a data-parallel step exchanges: the gradient (coarse model, fine
            # model) and, behind it, [loss, 0, 0, 0] -- one exchange per step
            self._grad_buf = torch.zeros(PARAMS_TOTAL + 4, device=self.device, dtype=torch.float32)
            self.flat_grads = self._grad_buf[:PARAMS_TOTAL]
            h = C.c_void_p()
            check(load().nerfb200_create(C.byref(h)), "create")
            self._ctx = h
            # operand images are packed only for the precisions this model uses (others are added on first use)
            self._pack_mask = 0
            for pz in (self.precision, self.train_precision):
                self._pack_mask |= _lib.PACK_BIT.get(pz, 0)
            check(load().nerfb200_set_option(h, _lib.OPT_PACK_MASK, self._pack_mask or 1), "set_option")
            check(load().nerfb200_set_option(h, _lib.OPT_PRECISE_LAST, int(self.precise_last)), "set_option")
        # SMs given to the coarse model's weight-gradient phase while the fine model's backward-data phase runs on the
        # others; 0 (default) = plain sequential backward. Measured on B200: 48 SMs 5.45 ms/step
        # vs 5.43 sequential, 32 and 64 worse - both phases slow down in proportion to the SMs they lose, so the overlap
        # buys nothing; kept as an option because the phase API is what a different schedule would build on.
        self._num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count if torch.cuda.is_available() else 148
        self._dw_overlap_sms = 0          # (an attribute, not an environment switch: tests set it to exercise the phase API)
        self._side_stream = None
        offs = _lib.param_offsets()
        self._variables = []
        for mi, mname in enumerate(("coarse", "fine")):
            names = variable_names(mname)
            for vi, nm in enumerate(names):
                ln = LAYER_NAMES[vi // 2]
                fi, fo = LAYER_SHAPES[ln]
                shape = (fi, fo) if vi % 2 == 0 else (fo,)
                o = mi * PARAMS_PER_MODEL + offs[vi]
                n = int(np.prod(shape))
                var = Variable(nm, self.flat_params[o:o + n].view(shape), owner=self)
                var._ofs, var._n = o, n
                self._variables.append(var)
        self.coarse_model = SubModel(self, COARSE, "coarse")
        self.fine_model = SubModel(self, FINE, "fine")
        self._dirty = True
        self._ws = {}
        self.optimizer = None
        self.metrics = []
        self.process_group = None
        self.world_size, self.rank = 1, 0
        self._peer_owner = None
        self._peer = None                 # nerfb200_peer handle: the gradient exchange over NVLink peer memory
        self.fuse_exchange_adam = True    # ... with the Adam step in the same launch
        self.peer_mode = None             # "nvls" (switch reduces/replicates), "symm-p2p" or "ipc" (unicast loads/stores)
        self.peer_symmetric_memory = True   # map the blocks through torch symmetric memory (else: the library's CUDA IPC)
        self.peer_multicast = None        # NVLS multicast mapping: None = from 4 ranks up, True / False = always / never
        self.overlap_allreduce = True     # NCCL path: all-reduce the coarse gradient while the fine backward runs
        self.graph_overlap_allreduce = True   # the same fork/join inside a captured step
        self.fused_forward = True         # forward()/predict()/render: the whole march as one C-ABI call (nerfb200_forward)
        self.use_cuda_graph = bool(int(cuda_graph))     # train_step as one CUDA graph per batch shape (after two eager steps)
        self._graphs, self._step_dev, self._step_dev_host = {}, None, None
        self._step_counter = 0
        self.last_loss = None

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                load().nerfb200_destroy(self._ctx)
                self._ctx = None
            # a peer block is NOT freed here: other ranks may still have it mapped (close_distributed does it in step)
        except Exception:
            pass

    # ------------------------------------------------------------------ Keras-shaped plumbing
    @property
    def trainable_variables(self):
        return self._variables

    def compile(self, optimizer=None, metrics=None, run_eagerly=False):
        """Model.compile as used by setup_model (core/model.py:422-427). The optimiser is always the
        reference's Adam + ExponentialDecay; `optimizer` is accepted for signature parity."""
        self.optimizer = _Adam(self)
        self.metrics = list(metrics) if metrics is not None else [ops.PSNRMetric()]

    def set_distributed(self, process_group=None, peer_exchange=True):
        """Data-parallel training over torch.distributed (one process per GPU): the 4096-ray batch is split across the
        ranks and the flat [gradient | loss] buffer is summed over the ranks once per step.
        `peer_exchange` (default): the ranks map each other's gradient block over NVLink -- through torch symmetric
        memory (with an NVLS multicast mapping on an NVSwitch box) or, failing that, the library's own CUDA-IPC
        mapping -- and ONE kernel of this library per rank does the exchange, fused with the Adam step (csrc/peer.cu,
        DESIGN.md 4.2c); NCCL then only carries the mapping handshake. Falls back to an NCCL all-reduce, with a warning,
        if the ranks cannot map each other's memory (more than 8 ranks, several nodes, no peer access).
        `peer_symmetric_memory` / `peer_multicast` (attributes, set before this call) select the mapping."""
        import torch.distributed as dist
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self.rank = dist.get_rank(self.process_group)
        if peer_exchange and self.world_size > 1 and self._peer is None:
            self._setup_peer_exchange()

    @_on_device
    def _setup_peer_exchange(self):
        """Maps the ranks' gradient blocks into each other. First choice: torch symmetric memory (plumbing: it allocates
        the block, exchanges the handles and, on an NVSwitch box, binds an NVLS multicast mapping) handed to
        nerfb200_peer_attach; second: the library's own CUDA-IPC mapping (nerfb200_peer_create/handle/connect, unicast
        loads and stores). Every decision is agreed on by all ranks before the next collective step."""
        import torch.distributed as dist
        import warnings
        lib, W, n = load(), self.world_size, PARAMS_TOTAL + 4

        def agreed(ok):
            flag = torch.tensor([int(bool(ok))], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.process_group)
            return bool(int(flag.item()))

        if W > 8:
            warnings.warn("peer-memory gradient exchange: more than 8 ranks; using the NCCL all-reduce")
            return
        h, buf, keep, mode, why = C.c_void_p(), None, None, None, ""
        # ---- 1. symmetric memory (+ NVLS multicast where the box has it)
        block = None
        if self.peer_symmetric_memory:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                block = symm_mem.empty(1024 + n, dtype=torch.float32, device=self.device)     # 4096 header bytes + floats
                block.zero_()
                torch.cuda.synchronize(self.device)
            except Exception as ex:
                block, why = None, f"symmetric memory: {type(ex).__name__}: {ex}"
        if agreed(block is not None):
            try:
                hdl = symm_mem.rendezvous(block, self.process_group)
                off = block.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])
                assert hdl.world_size == W and hdl.rank == self.rank and 0 <= off and off + block.numel() * 4 <= hdl.buffer_size
                bases = (C.c_void_p * W)(*[int(p) + off for p in hdl.buffer_ptrs])
                # measured (tools/dp_bench.py, 4.77 MB): 2 GPUs 23 us unicast / 31 us multicast, 8 GPUs 32 / 30 us
                use_mc = self.peer_multicast if self.peer_multicast is not None else W >= 4
                mc = int(hdl.multicast_ptr) + off if (use_mc and int(hdl.multicast_ptr)) else 0
                check(lib.nerfb200_peer_attach(W, self.rank, n, bases, C.c_void_p(mc), C.byref(h)), "peer_attach")
                buf, keep, mode = block[1024:], (block, hdl), ("nvls" if mc else "symm-p2p")
            except Exception as ex:
                h, buf, why = C.c_void_p(), None, f"symmetric memory: {type(ex).__name__}: {ex}"
            if agreed(buf is not None):
                dist.barrier(group=self.process_group)          # every header is zeroed and mapped before the first flag
            else:
                h, buf = C.c_void_p(), None
        # ---- 2. the library's own CUDA-IPC mapping
        if buf is None:
            handle = C.create_string_buffer(64)
            ok = lib.nerfb200_peer_create(W, self.rank, n, C.byref(h)) == 0 and lib.nerfb200_peer_handle(h, handle) == 0
            if not ok:
                why = lib.nerfb200_last_error().decode("utf-8", "replace")
            mine = torch.tensor(list(handle.raw) + [int(ok)], dtype=torch.uint8, device=self.device)
            gathered = [torch.empty_like(mine) for _ in range(W)]
            dist.all_gather(gathered, mine, group=self.process_group)       # the 64-byte handles go round by NCCL
            blob = torch.stack(gathered).cpu().numpy()
            if ok and bool(blob[:, 64].all()):
                if lib.nerfb200_peer_connect(h, blob[:, :64].tobytes()) != 0:
                    ok, why = False, lib.nerfb200_last_error().decode("utf-8", "replace")
                else:
                    try:
                        addr = C.c_void_p()
                        check(lib.nerfb200_peer_buffer(h, C.byref(addr)), "peer_buffer")
                        buf = torch.as_tensor(_DeviceFloats(addr.value, n), device=self.device)
                        assert buf.data_ptr() == addr.value and buf.dtype == torch.float32 and buf.numel() == n
                        keep, mode = buf, "ipc"
                    except Exception as ex:          # torch could not wrap the library's memory
                        ok, why, buf = False, f"{type(ex).__name__}: {ex}", None
            else:
                ok = False
            if not agreed(ok and buf is not None):
                # (the block, if any, stays allocated: a peer may have mapped it already)
                warnings.warn("peer-memory gradient exchange unavailable on some rank"
                              + (f" (this rank: {why})" if why else "") + "; using the NCCL all-reduce")
                return
        buf.copy_(self._grad_buf)
        self._peer, self._peer_owner, self.peer_mode = h, keep, mode
        self._grad_buf = buf
        self.flat_grads = self._grad_buf[:PARAMS_TOTAL]
        self._graphs.clear()              # captured steps hold the old gradient buffer

    def close_distributed(self):
        """Releases the peer mappings (collective: every rank calls it; nothing is freed while a peer may still use it)."""
        if self._peer is None:
            return
        import torch.distributed as dist
        self.release_cuda_graphs()
        dist.barrier(group=self.process_group)
        with torch.cuda.device(self.device):
            old = self._grad_buf
            if self.last_loss is not None:
                self.last_loss = self.last_loss.clone()
            self._grad_buf = old.clone()
            self.flat_grads = self._grad_buf[:PARAMS_TOTAL]
            torch.cuda.synchronize(self.device)
            del old
            self._peer_owner = None
            check(load().nerfb200_peer_disconnect(self._peer), "peer_disconnect")
            dist.barrier(group=self.process_group)
            check(load().nerfb200_peer_destroy(self._peer), "peer_destroy")
        self.peer_mode = None
        self._peer = None

    def set_flat_params(self, flat):
        self.flat_params.copy_(torch.as_tensor(flat, dtype=torch.float32).to(self.device))
        self._dirty = True

    def set_weights_from_dict(self, w):
        """Load a {name: array} dict keyed like variable_names()."""
        for var in self._variables:
            var.assign(w[var.name])
        self._dirty = True

    def set_everything(self, load_dir=None, load_tag=None, skip_optimizer=None):
        """NeRF.set_everything (core/model.py:239-287); directory, tag and skip_optimizer default to params.model.load.*"""
        from . import checkpoint
        load = getattr(getattr(self.params, "model", None), "load", None)
        load_dir = load_dir if load_dir is not None else load.load_dir
        load_tag = load_tag if load_tag is not None else load.load_tag
        if skip_optimizer is None:
            skip_optimizer = bool(getattr(load, "skip_optimizer", False)) if load is not None else False
        checkpoint.set_everything(self, load_dir, load_tag, skip_optimizer)

    def _sync_packed(self, precision=None):
        bit = _lib.PACK_BIT.get(precision, 0)
        if bit and not (self._pack_mask & bit):          # a precision this model has not used before
            self._pack_mask |= bit
            check(load().nerfb200_set_option(self._ctx, _lib.OPT_PACK_MASK, self._pack_mask), "set_option")
            self._dirty = True
        if self._dirty:
            check(load().nerfb200_pack_weights(self._ctx, ptr(self.flat_params), stream_ptr()), "pack_weights")
            self._dirty = False

    def _scratch(self, key, nbytes):
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 16), device=self.device, dtype=torch.uint8)
            self._ws[key] = buf
        return buf

    # ------------------------------------------------------------------------------ MLP calls
    def _mlp(self, which, rays_o, rays_d, t_vals, precision=None, stash=None):
        precision = self.precision if precision is None else precision
        B, S = t_vals.shape
        R = B * S
        rgb = torch.empty((R, 3), device=self.device, dtype=torch.float32)
        sigma = torch.empty((R,), device=self.device, dtype=torch.float32)
        self._sync_packed(precision)
        training = stash is not None
        wsb = load().nerfb200_mlp_workspace_bytes(R, precision, 0)
        ws = self._scratch("mlp_ws", wsb) if (wsb > 0 and not training) else None
        check(load().nerfb200_mlp_forward(self._ctx, which, B, S, ptr(rays_o), ptr(rays_d), ptr(t_vals.contiguous()),
                                          ptr(self.flat_params), ptr(rgb), ptr(sigma), precision,
                                          ptr(ws, torch.uint8, allow_none=True),
                                          ptr(stash, torch.uint8, allow_none=True), stream_ptr()), "mlp_forward")
        return rgb, sigma

    def _mlp_backward(self, which, rays_o, rays_d, t_vals, d_rgb, d_sigma, precision, stash):
        B, S = t_vals.shape
        wsb = load().nerfb200_mlp_workspace_bytes(B * S, precision, 1)
        ws = self._scratch("mlp_bwd_ws", wsb) if wsb > 0 else None
        check(load().nerfb200_mlp_backward(self._ctx, which, B, S, ptr(rays_o), ptr(rays_d), ptr(t_vals.contiguous()),
                                           ptr(self.flat_params), ptr(d_rgb), ptr(d_sigma), ptr(self.flat_grads),
                                           precision, ptr(ws, torch.uint8, allow_none=True),
                                           ptr(stash, torch.uint8, allow_none=True), stream_ptr()), "mlp_backward")

    # ---------------------------------------------------------------------------- the ray march
    @_on_device
    def forward(self, rays_o, rays_d, near, far, u_coarse=None, u_fine=None, ray0=0, precision=None,
                need_weights=True, _train=None, _step_state=None):
        """NeRF.forward (core/model.py:57-125): stratified sampling -> coarse MLP -> compositing ->
        hierarchical sampling -> fine MLP -> compositing. Returns (post_proc_CM, post_proc_FM).
        `u_coarse`/`u_fine` are the fixed uniforms of the parity contract; when None the kernels
        draw Philox uniforms keyed by (rng_seed, step, ray0 + ray)."""
        s = self.params.sampling
        precision = self.precision if precision is None else precision
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        # the sampling noise is keyed by (rng_seed, step, global ray id); with a device-resident step state the step is
        # XORed in on the device, so that a captured step draws fresh noise at every replay
        seed = (self.rng_seed << 20) ^ (self._step_counter if _step_state is None else 0)
        if _train is None and self.fused_forward and precision != FP32:
            # the whole march as ONE C-ABI call (nerfb200_forward): six to eight launches back to back, intermediates in a
            # persistent workspace. (With need_weights=False the coarse weights stay in the workspace too.)
            return self._forward_one_call(rays_o, rays_d, near, far, u_coarse, u_fine, seed, _step_state, ray0, precision, need_weights)
        t_c, edges = ray_utils.sample_coarse(s.N_coarse, s.lin_inv_depth, s.perturb, near, far, u_coarse, seed, ray0,
                                             step_state=_step_state)
        st_c = st_f = None
        if _train is not None:
            R_c, R_f = t_c.shape[0] * s.N_coarse, t_c.shape[0] * (s.N_coarse + s.N_fine)
            st_c = self._scratch("stash_c", load().nerfb200_mlp_stash_bytes(R_c, precision))
            st_f = self._scratch("stash_f", load().nerfb200_mlp_stash_bytes(R_f, precision))
        rgb_c, sig_c = self._mlp(COARSE, rays_o, rays_d, t_c, precision, st_c)
        if _train is not None:
            # training: integrator + loss term + integrator backward of each model as ONE launch (nerfb200_composite_train)
            pp_c, ds_c, dr_c = ray_utils.composite_train(rgb_c, sig_c, t_c, self.white_bg, _train["gt"], _train["Bg"], _train["loss"])
        else:
            pp_c = ray_utils.post_process_model_output(rgb_c, sig_c, t_c, self.white_bg)
        t_f = ray_utils.sample_fine(s.N_fine, pp_c["weights"], edges, t_c, u_fine, seed, ray0, step_state=_step_state)
        rgb_f, sig_f = self._mlp(FINE, rays_o, rays_d, t_f, precision, st_f)
        if _train is not None:
            pp_f, ds_f, dr_f = ray_utils.composite_train(rgb_f, sig_f, t_f, self.white_bg, _train["gt"], _train["Bg"], _train["loss"],
                                                         metric_state=_train["metric"], need_weights=False)
            _train.update(dict(t_c=t_c, t_f=t_f, st_c=st_c, st_f=st_f, ds_c=ds_c, dr_c=dr_c, ds_f=ds_f, dr_f=dr_f))
        else:
            pp_f = ray_utils.post_process_model_output(rgb_f, sig_f, t_f, self.white_bg, need_weights=need_weights)
        return pp_c, pp_f

    def _forward_one_call(self, rays_o, rays_d, near, far, u_coarse, u_fine, seed, step_state, ray0, precision, need_weights):
        s = self.params.sampling
        lib = load()
        B, Nc, Nf = int(rays_o.shape[0]), int(s.N_coarse), int(s.N_fine)
        f32 = dict(device=self.device, dtype=torch.float32)
        pp_c = {"acc_map": torch.empty((B,), **f32), "pred_rgb": torch.empty((B, 3), **f32), "pred_depth": torch.empty((B,), **f32)}
        pp_f = {"acc_map": torch.empty((B,), **f32), "pred_rgb": torch.empty((B, 3), **f32), "pred_depth": torch.empty((B,), **f32)}
        if need_weights:
            pp_c["weights"] = torch.empty((B, Nc), **f32)
            pp_f["weights"] = torch.empty((B, Nc + Nf), **f32)
        if B == 0:
            return pp_c, pp_f
        self._sync_packed(precision)
        ws = self._scratch("fwd_ws", lib.nerfb200_forward_workspace_bytes(B, Nc, Nf))
        opt = lambda t: ptr(t.contiguous(), allow_none=True) if t is not None else C.c_void_p(0)
        check(lib.nerfb200_forward(self._ctx, B, Nc, Nf, int(bool(s.lin_inv_depth)), int(bool(s.perturb)), int(bool(self.white_bg)),
                                   ptr(rays_o), ptr(rays_d), ptr(near.reshape(-1).contiguous()), ptr(far.reshape(-1).contiguous()),
                                   opt(u_coarse), opt(u_fine), seed, ptr(step_state, torch.int64, allow_none=True), ray0,
                                   ptr(self.flat_params), precision, ptr(ws, torch.uint8),
                                   ptr(pp_c["pred_rgb"]), ptr(pp_c["pred_depth"]), ptr(pp_c["acc_map"]), opt(pp_c.get("weights")),
                                   ptr(pp_f["pred_rgb"]), ptr(pp_f["pred_depth"]), ptr(pp_f["acc_map"]), opt(pp_f.get("weights")),
                                   stream_ptr()), "forward")
        return pp_c, pp_f

    def call(self, inputs):
        """NeRF.call (core/model.py:36-55): inputs[0] = (rays_o, rays_d, near, far)."""
        ro, rd, near, far = (self._to_device(a) for a in inputs[0])
        return self.forward(ro, rd, near, far)

    __call__ = call

    def _to_device(self, a):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.to(dtype=torch.float32).contiguous()
            return a.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    # -------------------------------------------------------------------------------- training
    def _loss_and_grads(self, rays_o, rays_d, near, far, rgb, u_coarse=None, u_fine=None, ray0=0,
                        global_batch=None, coarse_done=None, step_state=None):
        """Forward + backward of train_step (core/model.py:148-170) into self.flat_grads (local sum).
        Returns the device tensor [loss] (this rank's share of the global mean losses)."""
        prec = self.train_precision
        B = rays_o.shape[0]
        Bg = B * self.world_size if global_batch is None else global_batch
        self._grad_buf.zero_()                       # the kernels accumulate into the gradient and the loss
        loss = self._grad_buf[PARAMS_TOTAL:PARAMS_TOTAL + 1]
        metric = self.metrics[0] if self.metrics else None
        if metric is not None:
            metric._ensure(self.device)
        # the forward's integrators also add the two MeanSquaredError terms to `loss` (and the fine one to the metric
        # state) and leave d(loss)/d(sigma, rgb) of both models in `tr` (ray_utils.composite_train)
        tr = {"gt": rgb.contiguous(), "Bg": Bg, "loss": loss, "metric": metric.state if metric is not None else None}
        pp_c, pp_f = self.forward(rays_o, rays_d, near, far, u_coarse, u_fine, ray0, precision=prec, _train=tr,
                                  _step_state=step_state)
        lib = load()
        ds_c, dr_c, ds_f, dr_f = tr["ds_c"], tr["dr_c"], tr["ds_f"], tr["dr_f"]
        n_dw = self._dw_overlap_sms if prec != FP32 else 0
        if n_dw <= 0:
            self._mlp_backward(COARSE, rays_o, rays_d, tr["t_c"], dr_c, ds_c, prec, tr["st_c"])
            if coarse_done is not None:
                coarse_done()                        # data-parallel: the coarse half of the gradient is final
            self._mlp_backward(FINE, rays_o, rays_d, tr["t_f"], dr_f, ds_f, prec, tr["st_f"])
            return loss, pp_c, pp_f
        # Phase-split backward. Per model: backward-data (writes the gradient stash: HBM-write bound), then the
        # weight-gradient GEMM (reads both stashes: HBM-read bound). The coarse model's weight-gradient phase runs on a
        # second stream, on `n_dw` SMs, NEXT TO the fine model's backward-data phase on the remaining SMs, so reads and
        # writes are in flight together; the two models write disjoint halves of flat_grads.
        (Bc, Sc), (Bf, Sf) = tr["t_c"].shape, tr["t_f"].shape
        ws_c = self._scratch("mlp_bwd_ws_c", lib.nerfb200_mlp_workspace_bytes(Bc * Sc, prec, 1))
        ws_f = self._scratch("mlp_bwd_ws_f", lib.nerfb200_mlp_workspace_bytes(Bf * Sf, prec, 1))
        main = torch.cuda.current_stream(self.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        u8 = torch.uint8
        check(lib.nerfb200_mlp_backward_data(self._ctx, COARSE, Bc, Sc, ptr(self.flat_params), ptr(dr_c), ptr(ds_c), prec,
                                             ptr(ws_c, u8), ptr(tr["st_c"], u8), 0, stream_ptr()), "mlp_backward_data")
        ev = torch.cuda.Event()
        ev.record(main)
        check(lib.nerfb200_mlp_backward_data(self._ctx, FINE, Bf, Sf, ptr(self.flat_params), ptr(dr_f), ptr(ds_f), prec,
                                             ptr(ws_f, u8), ptr(tr["st_f"], u8), self._num_sms - n_dw, stream_ptr()),
              "mlp_backward_data")
        side.wait_event(ev)
        check(lib.nerfb200_mlp_backward_weights(self._ctx, COARSE, Bc, Sc, ptr(self.flat_grads), prec, ptr(ws_c, u8),
                                                ptr(tr["st_c"], u8), n_dw, C.c_void_p(side.cuda_stream)),
              "mlp_backward_weights")
        main.wait_stream(side)
        check(lib.nerfb200_mlp_backward_weights(self._ctx, FINE, Bf, Sf, ptr(self.flat_grads), prec, ptr(ws_f, u8),
                                                ptr(tr["st_f"], u8), 0, stream_ptr()), "mlp_backward_weights")
        return loss, pp_c, pp_f

    @_on_device
    def train_step(self, data, u_coarse=None, u_fine=None, ray0=None):
        """NeRF.train_step (core/model.py:127-180). `data` = ((rays_o, rays_d, near, far), (rgb,)) --
        this rank's shard of the batch when data-parallel. The in-kernel Philox streams are keyed by
        (seed, step, ray0 + ray); `ray0` defaults to rank * B so that ranks draw independent noise."""
        if self.optimizer is None:
            self.compile()
        (ro, rd, near, far), (rgb,) = data
        ro, rd, near, far, rgb = (self._to_device(a) for a in (ro, rd, near, far, rgb))
        if ray0 is None:
            ray0 = self.rank * int(ro.shape[0])
        if self.use_cuda_graph and u_coarse is None and u_fine is None and self._dw_overlap_sms <= 0:
            if self._train_step_graphed(ro, rd, near, far, rgb, ray0):
                return {m.name: m.result_async() for m in self.metrics}
        pending = []
        coarse_done = None
        if self.world_size > 1 and self.overlap_allreduce and self._peer is None:
            import torch.distributed as dist

            def coarse_done():
                # the coarse model's gradient is final before the fine model's backward starts: its all-reduce runs on
                # NCCL's stream next to the fine backward (it takes SMs as the persistent kernels' CTAs retire)
                pending.append(dist.all_reduce(self._grad_buf[:PARAMS_PER_MODEL], group=self.process_group, async_op=True))
        loss, _, _ = self._loss_and_grads(ro, rd, near, far, rgb, u_coarse, u_fine, ray0, coarse_done=coarse_done)
        self._exchange_and_apply(pending)
        self._step_counter += 1
        self.last_loss = loss
        return {m.name: m.result_async() for m in self.metrics}     # no device sync: see PSNRMetric.result_async

    def _exchange_and_apply(self, pending, step_state=None):
        """Sum of the flat [gradient | loss] buffer over the ranks, then Adam (core/model.py:170-171)."""
        if self.world_size > 1:
            if self._peer is not None:
                if self.fuse_exchange_adam:       # ONE launch: exchange over NVLink peer memory + Adam
                    self.optimizer.exchange_and_apply(self._peer, step_state=step_state)
                    return
                check(load().nerfb200_peer_allreduce(self._peer, stream_ptr()), "peer_allreduce")
            else:
                import torch.distributed as dist
                if pending:       # fine half + [loss] tail; then both must have landed before Adam reads them
                    pending.append(dist.all_reduce(self._grad_buf[PARAMS_PER_MODEL:], group=self.process_group, async_op=True))
                    for h in pending:
                        h.wait()
                else:
                    dist.all_reduce(self._grad_buf, group=self.process_group)     # ONE collective: gradient + loss
        self.optimizer.apply_gradients(self.flat_grads, step_state=step_state)

    @_on_device
    def _train_step_graphed(self, ro, rd, near, far, rgb, ray0):
        """The whole step -- sampling, both forwards, loss, both backwards, the gradient all-reduce, Adam, the repack of
        the operand images -- as ONE CUDA graph per batch shape: ~30 launches per step are otherwise issued one by
        one from Python, which is what bounds a data-parallel step of 512 rays per GPU. The first two steps of a
        shape run eagerly (they size every scratch buffer), the third is captured, later ones replay it. What changes
        from step to step lives on the device: the input batch (copied into static buffers), the sampling step and
        the optimizer iteration (`_step_dev`, advanced inside the graph). Returns False when the step must run eagerly."""
        for m in self.metrics:
            m._ensure(self.device)
        # a graph is tied to the batch shape, the ray offset and the buffers it accumulates the metric into
        key = (int(ro.shape[0]), int(ray0), tuple(m.state.data_ptr() for m in self.metrics))
        st = self._graphs.setdefault(key, {"calls": 0})
        st["calls"] += 1
        if st.get("failed") or st["calls"] <= 2:
            return False
        lib = load()
        if self._step_dev is None:
            self._step_dev = torch.zeros(2, dtype=torch.int64, device=self.device)
            self._step_dev_host = None
        want = (int(self.optimizer.iterations), int(self._step_counter))
        if self._step_dev_host != want:                       # eager steps ran in between: re-seed the device counters
            self._step_dev.copy_(torch.tensor(want, dtype=torch.int64))
            self._step_dev_host = want
        if "graph" not in st:
            static = [torch.empty_like(x) for x in (ro, rd, near, far, rgb)]
            self._sync_packed(self.train_precision)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    pending, coarse_done = [], None
                    if self.world_size > 1 and self.overlap_allreduce and self.graph_overlap_allreduce and self._peer is None:
                        import torch.distributed as dist

                        def coarse_done():      # a fork inside the capture: NCCL's stream joins again at the wait() below
                            pending.append(dist.all_reduce(self._grad_buf[:PARAMS_PER_MODEL], group=self.process_group, async_op=True))
                    self._loss_and_grads(*static, ray0=ray0, step_state=self._step_dev, coarse_done=coarse_done)
                    self._exchange_and_apply(pending, step_state=self._step_dev)
                    check(lib.nerfb200_step_advance(ptr(self._step_dev, torch.int64), stream_ptr()), "step_advance")
                    check(lib.nerfb200_pack_weights(self._ctx, ptr(self.flat_params), stream_ptr()), "pack_weights")
            except Exception as ex:       # e.g. a collective that cannot be captured: stay eager, loudly
                import warnings
                warnings.warn(f"CUDA-graph capture of train_step failed ({ex}); continuing with eager launches")
                st["failed"] = True
                torch.cuda.synchronize(self.device)
                return False
            st["graph"], st["static"] = g, static
        for dst, src in zip(st["static"], (ro, rd, near, far, rgb)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._sync_packed(self.train_precision)       # weights assigned from outside since the last step (the graph repacks after Adam)
        st["graph"].replay()
        self.optimizer.iterations += 1
        self._step_counter += 1
        self._step_dev_host = (int(self.optimizer.iterations), int(self._step_counter))
        self._dirty = False                                   # the graph repacked the operand images after Adam
        self.last_loss = self._grad_buf[PARAMS_TOTAL:PARAMS_TOTAL + 1]
        return True

    def release_cuda_graphs(self):
        """Drops the captured training-step graphs (they hold their private memory pool and, data-parallel, captured NCCL
        work: release them before destroying the process group)."""
        torch.cuda.synchronize(self.device)
        self._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize(self.device)

    @_on_device
    def test_step(self, data, u_coarse=None, u_fine=None, ray0=0):
        """NeRF.test_step (core/model.py:182-223): forward + metric update on the fine output."""
        (ro, rd, near, far), (rgb,) = data
        ro, rd, near, far, rgb = (self._to_device(a) for a in (ro, rd, near, far, rgb))
        _, pp_f = self.forward(ro, rd, near, far, u_coarse, u_fine, ray0, need_weights=False)
        for m in self.metrics:
            m.update_state(rgb, pp_f["pred_rgb"])
        return {m.name: m.result_async() for m in self.metrics}

    def predict_step(self, data):
        """NeRF.predict_step (core/model.py:225-237)."""
        return self(data)

    def _metric_allreduce(self):
        if self.world_size > 1:
            import torch.distributed as dist
            for m in self.metrics:
                if m.state is not None:
                    dist.all_reduce(m.state, group=self.process_group)

    def fit(self, x=None, epochs=1, steps_per_epoch=None, validation_data=None, validation_freq=1,
            callbacks=None, initial_epoch=0, verbose=0, **_):
        """Model.fit as driven by main/train.py:57-64: `steps_per_epoch` train_steps per epoch over
        the (repeating) dataset, metrics reset per epoch, validation every `validation_freq` epochs."""
        if self.optimizer is None:
            self.compile()
        hist = History()
        it = iter(x)
        for cb in callbacks or []:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        for epoch in range(initial_epoch, epochs):
            for m in self.metrics:
                m.reset_states()
            steps, logs = 0, {}
            while steps_per_epoch is None or steps < steps_per_epoch:
                try:
                    batch = next(it)
                except StopIteration:
                    if steps_per_epoch is None:
                        break
                    it = iter(x)
                    batch = next(it)
                self.train_step(batch)
                steps += 1
            self._metric_allreduce()
            logs = {m.name: m.result() for m in self.metrics}
            if self.last_loss is not None:
                logs["loss"] = float(self.last_loss.item())
            if validation_data is not None and (epoch + 1) % validation_freq == 0:
                val = self.evaluate(validation_data, return_dict=True)
                logs.update({f"val_{k}": v for k, v in val.items()})
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            for cb in callbacks or []:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if verbose:
                print(f"epoch {epoch + 1}/{epochs} " + " ".join(f"{k}={v:.4f}" for k, v in logs.items()))
            if steps_per_epoch is None:
                it = iter(x)
        return hist

    def evaluate(self, x=None, return_dict=False, **_):
        """Model.evaluate: test_step over the dataset; returns the PSNRMetric value."""
        if not self.metrics:
            self.compile()
        for m in self.metrics:
            m.reset_states()
        ray0 = 0
        for batch in x:
            self.test_step(batch, ray0=ray0)         # every validation batch draws its own sampling noise
            ray0 += int(batch[0][0].shape[0])
        self._metric_allreduce()
        res = {m.name: m.result() for m in self.metrics}
        return res if return_dict else (list(res.values())[0] if len(res) == 1 else list(res.values()))

    # ------------------------------------------------------------------------------- rendering
    @_on_device
    def render_rays(self, rays_o, rays_d, near, far, ray0=0, need_weights=False, u_coarse=None, u_fine=None,
                    keep_coarse=True):
        """Ray-march an arbitrary number of device-resident rays in chunks of `render_chunk`. Returns
        (dict_CM, dict_FM) of device tensors concatenated over chunks."""
        N = rays_o.shape[0]
        outs_c, outs_f = [], []
        for s0 in range(0, max(N, 1), self.render_chunk):
            s1 = min(N, s0 + self.render_chunk)
            if s1 <= s0:
                break
            uc = None if u_coarse is None else u_coarse[s0:s1].contiguous()
            uf = None if u_fine is None else u_fine[s0:s1].contiguous()
            pc, pf = self.forward(rays_o[s0:s1], rays_d[s0:s1], near[s0:s1], far[s0:s1], uc, uf, ray0 + s0,
                                  need_weights=need_weights)
            if not need_weights:
                pc.pop("weights", None)
            if keep_coarse:
                outs_c.append(pc)
            outs_f.append(pf)
        cat = lambda ds: {k: torch.cat([d[k] for d in ds], dim=0) for k in ds[0]} if ds else {}
        return cat(outs_c), cat(outs_f)

    @_on_device
    def predict(self, x=None, return_weights=True, as_numpy=True, **_):
        """Model.predict over a render dataset (main/eval.py:48, main/render.py:85): returns
        (dict_CM, dict_FM) with keys acc_map [N], weights [N,S], pred_rgb [N,3], pred_depth [N],
        concatenated over the dataset's batches, as NumPy arrays like Keras does. Batches are
        grouped into super-chunks of `render_chunk` rays before they hit the GPU (rays are
        independent, so the result does not depend on the grouping). `return_weights=False` drops
        the per-sample `weights` tensors (0.66 GB per 800x800 view that no reference consumer reads)."""
        group, count = [], 0
        res_c, res_f = [], []            # per chunk: dict of device tensors, or of pinned host tensors (as_numpy)
        ray0, chunk_idx = 0, 0
        pending = []                     # the chunk whose D2H copies are still in flight
        keep = []                        # device tensors with a D2H copy in flight
        cur = torch.cuda.current_stream(self.device)
        cs = self._ws.get("copy_stream")
        if cs is None:
            cs = self._ws["copy_stream"] = torch.cuda.Stream(device=self.device)

        def pinned(tag, shape, dtype=torch.float32):
            """Persistent pinned staging buffers (cudaHostAlloc per call would cost more than the copy)."""
            key = ("pin", tag, tuple(shape), dtype)
            buf = self._ws.get(key)
            if buf is None:
                buf = self._ws[key] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            return buf

        def stage_in(cols):
            """Host batches -> one pinned buffer per ray field (double-buffered) -> device, on the copy stream."""
            slot = chunk_idx & 1
            ev = self._ws.get(("in_done", slot))
            if ev is not None:
                ev.synchronize()                      # the previous H2D out of this staging slot has finished
            dev = []
            cap = -(-count // 65536) * 65536
            with torch.cuda.stream(cs):
                for f, col in enumerate(cols):
                    width = int(np.asarray(col[0]).reshape(col[0].shape[0], -1).shape[1])
                    pin = pinned(("in", slot, f), (cap, width))
                    n = 0
                    view = pin.numpy()
                    for c in col:
                        c = np.asarray(c, dtype=np.float32).reshape(c.shape[0], -1)
                        view[n:n + c.shape[0]] = c
                        n += c.shape[0]
                    d = pin[:n].to(self.device, non_blocking=True)
                    d.record_stream(cur)
                    dev.append(d)
                ev = self._ws[("in_done", slot)] = torch.cuda.Event()
                ev.record(cs)
            cur.wait_event(ev)
            return dev

        def stage_out(which, d):
            """Device results of one chunk -> pinned host buffers, on the copy stream, overlapping the next chunk."""
            done = torch.cuda.Event()
            done.record(cur)
            out = {}
            with torch.cuda.stream(cs):
                cs.wait_event(done)
                for k, v in d.items():
                    # two staging slots (the per-sample `weights` [N,S] of the Keras-default return included: 50 MB per
                    # 65536-ray fine chunk, streamed like everything else instead of one blocking 0.66 GB copy per view): chunk i-2 has been drained before chunk i is staged (see flush)
                    pin = pinned(("out", which, chunk_idx & 1, k), (-(-v.shape[0] // 65536) * 65536,) + tuple(v.shape[1:]))
                    pin[:v.shape[0]].copy_(v, non_blocking=True)
                    v.record_stream(cs)
                    out[k] = pin[:v.shape[0]]
                out["_done"] = torch.cuda.Event()
                out["_done"].record(cs)
            keep.append(d)
            return out

        # a plain (non-repeating, unskipped) RayDataset knows its length: results are written in place
        total = x.n if isinstance(x, RayDataset) and not x._repeat and not x._skip and not x.drop_remainder else None
        final = [{}, {}]

        def drain(which, out, row0):
            """Pinned chunk results -> the returned NumPy arrays (host work that overlaps the next chunk on the GPU)."""
            out.pop("_done").synchronize()
            res = {}
            for k, v in out.items():
                if v.is_cuda:
                    res[k] = v
                elif total is not None:
                    if k not in final[which]:
                        final[which][k] = np.empty((total,) + tuple(v.shape[1:]), dtype=np.float32)
                    final[which][k][row0:row0 + v.shape[0]] = v.numpy()
                else:
                    res[k] = torch.from_numpy(v.numpy().copy())
            return res

        def flush():
            nonlocal group, count, ray0, chunk_idx
            if not group:
                return
            cols = list(zip(*group))
            on_dev = isinstance(cols[0][0], torch.Tensor) and cols[0][0].is_cuda
            if on_dev:
                dev = [torch.cat(col, dim=0) if len(col) > 1 else col[0] for col in cols]
            else:
                dev = stage_in(cols)
            pc, pf = self.render_rays(dev[0], dev[1], dev[2].reshape(-1, 1), dev[3].reshape(-1, 1), ray0,
                                      need_weights=return_weights)
            if as_numpy:
                pc, pf = stage_out(0, pc), stage_out(1, pf)
                if pending:                               # the previous chunk's copies, while this one computes
                    qc, qf, r0 = pending.pop()
                    res_c.append(drain(0, qc, r0))
                    res_f.append(drain(1, qf, r0))
                pending.append((pc, pf, ray0))
            else:
                res_c.append(pc)
                res_f.append(pf)
            ray0 += count
            chunk_idx += 1
            group, count = [], 0

        for batch in x:
            inp = batch[0]
            group.append(inp)
            count += int(inp[0].shape[0])
            if count >= self.render_chunk:
                flush()
        flush()

        if as_numpy and pending:
            qc, qf, r0 = pending.pop()
            res_c.append(drain(0, qc, r0))
            res_f.append(drain(1, qf, r0))

        def finish(which, ds):
            if not ds:
                return {}
            if not as_numpy:
                return {k: torch.cat([d[k] for d in ds], dim=0) for k in ds[0]}
            out = dict(final[which])
            for k in ds[0]:                               # what was not written in place
                out[k] = np.concatenate([d[k].cpu().numpy() for d in ds], axis=0)
            return out

        return finish(0, res_c), finish(1, res_f)


def get_coarse_or_fine_model(model_name, num_units=256, params=None, **kw):
    """get_coarse_or_fine_model (core/model.py:334-394): returns the named sub-model of a fresh NeRF."""
    assert model_name in ("coarse", "fine")
    assert num_units == 256, "the reference hard-codes 256 units (core/model.py:334)"
    from .params import make_params
    nerf = NeRF(params or make_params(), **kw)
    return nerf.coarse_model if model_name == "coarse" else nerf.fine_model


def setup_model(params, **kw):
    """setup_model (core/model.py:396-432): Adam(ExponentialDecay(5e-4, 500000, 0.1)) + PSNRMetric; weights and
    optimiser state are restored from params.model.load.* when `set_weights` is on. `params.system.tf_seed`
    (main/train.py:20; may be None) seeds the sampling noise unless `rng_seed` is given."""
    if "rng_seed" not in kw:
        tf_seed = getattr(getattr(params, "system", None), "tf_seed", None)
        if tf_seed is not None:
            kw["rng_seed"] = int(tf_seed)
    nerf = NeRF(params=params, **kw)
    nerf.compile(optimizer="adam", metrics=[ops.PSNRMetric()],
                 run_eagerly=getattr(params.system, "run_eagerly", False))
    load = getattr(getattr(params, "model", None), "load", None)
    if load is not None and getattr(load, "set_weights", False):
        nerf.set_everything()
    return nerf


def setup_model_and_callbacks(params, num_imgs=None, img_HW=None, **kw):
    """setup_model_and_callbacks (core/model.py:434-483): the CustomSaver callback + the model. The reference's
    TensorBoard and LogValImages callbacks are logging only and not provided (out of scope)."""
    from .checkpoint import CustomSaver
    callbacks = [CustomSaver(params=params, save_best_only=False)]
    return setup_model(params, **kw), callbacks
