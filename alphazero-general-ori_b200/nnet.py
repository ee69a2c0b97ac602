"""Batched leaf evaluator: the inference half of the reference's network wrapper on the device.

Replaces `GenericNNetWrapper.predict` (GenericNNetWrapper.py:141-168, batch 1, CPU torch / ONNX) and
`SplendorNNet.forward` (SplendorNNet.py:127-159) for inference. Takes the tree arena's leaf rows as they are
(int8[B,R,7] states, uint8[B,406] legal masks, device tensors, no copy) and returns `exp(log_softmax(masked pi))` and
`tanh(v)` - exactly what `predict` hands to MCTS - as float32[B,406] / float32[B,n] device tensors.

Weights use the reference's own `state_dict` key names and shapes (so a checkpoint saved by
GenericNNetWrapper.save_checkpoint :185-198 loads), BatchNorm layers are folded for inference (eval mode, running
statistics), dropout is inactive at inference (SplendorNNet.py:133-140 with training=False). The score-difference head
is not evaluated: MCTS never reads it (GenericNNetWrapper.py:165).

Tensor cores appear only here (bf16 / tf32 matmuls through torch); the rules and tree kernels are integer code.
"""
import math

import numpy as np
import torch

NUM_ACTIONS = 406
BN_EPS = 1e-5


def _rows(n):
    return 32 + 10 * n + n * n


def state_dict_shapes(n_players):
    """key -> shape of the reference network's parameters and buffers (SplendorNNet.py:56-120)"""
    R, n = _rows(n_players), n_players
    sh = {}

    def lin(name, o, i):
        sh[name + ".weight"] = (o, i); sh[name + ".bias"] = (o,)

    def bn(name, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            sh[f"{name}.{k}"] = (c,)
        sh[name + ".num_batches_tracked"] = ()

    lin("dense2d_1.0", 128, R); bn("dense2d_1.1", 7); lin("dense2d_1.3", 128, 128)
    lin("partialgpool_1.dense_part.0", 120, 96); bn("partialgpool_1.dense_part.1", 7)
    lin("dense2d_3.0", 128, 128)
    lin("dense1d_4.0", 128, 704)
    lin("partialgpool_4.dense_part.0", 120, 112); bn("partialgpool_4.dense_part.1", 1)
    lin("dense1d_5.0", 128, 128); bn("dense1d_5.1", 1); lin("dense1d_5.3", 128, 128)
    lin("partialgpool_5.dense_part.0", 120, 112); bn("partialgpool_5.dense_part.1", 1)
    lin("output_layers_PI.0", 128, 128); lin("output_layers_PI.1", NUM_ACTIONS, 128)
    lin("output_layers_V.0", 128, 128); lin("output_layers_V.1", n, 128)
    lin("output_layers_SDIFF.0", 128, 128); lin("output_layers_SDIFF.1", n * 31, 128)
    sh["lowvalue"] = (1,)
    return sh


def random_state_dict(n_players, seed=0, trained_like=True):
    """Random-init weights of the reference architecture (kaiming-uniform weights as in SplendorNNet.py:67-73). With
    `trained_like` the biases and BatchNorm statistics are non-trivial too, so every folded term is exercised."""
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for k, shape in state_dict_shapes(n_players).items():
        if k == "lowvalue":
            sd[k] = torch.tensor([-1e8])
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100 if trained_like else 0)
        elif k.endswith("running_var"):
            sd[k] = (0.5 + torch.rand(shape, generator=g)) if trained_like else torch.ones(shape)
        elif k.endswith("running_mean"):
            sd[k] = (0.2 * torch.randn(shape, generator=g)) if trained_like else torch.zeros(shape)
        elif len(shape) == 2:
            bound = math.sqrt(6.0 / shape[1])   # kaiming_uniform_(a=0): gain sqrt(2), bound = gain * sqrt(3 / fan_in)
            sd[k] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif ".1.weight" in k and len(shape) == 1 and shape[0] in (1, 7):   # BatchNorm gamma
            sd[k] = (0.8 + 0.4 * torch.rand(shape, generator=g)) if trained_like else torch.ones(shape)
        else:   # biases (Linear and BatchNorm beta)
            sd[k] = (0.1 * torch.randn(shape, generator=g)) if trained_like else torch.zeros(shape)
    if trained_like:   # keep tanh(v) and the policy away from saturation, as a trained network's are
        sd["output_layers_V.1.weight"] *= 0.02
        sd["output_layers_PI.1.weight"] *= 0.05
    return sd


def fold(sd, device, dtype=torch.float32):
    """state_dict -> the tensors the inference pass uses (BatchNorm folded in float64, matrices pre-transposed, all in `dtype`)"""
    f = {k: v.detach().to(torch.float64) for k, v in sd.items() if torch.is_tensor(v) and v.dtype.is_floating_point}

    def bn(name):
        s = f[name + ".weight"] / torch.sqrt(f[name + ".running_var"] + BN_EPS)
        return s, f[name + ".bias"] - f[name + ".running_mean"] * s

    def lin(name):
        return f[name + ".weight"], f[name + ".bias"]

    W = {}

    def put(key, w, b):
        W[key + "_w"] = w.t().contiguous().to(device=device, dtype=dtype)
        W[key + "_b"] = b.to(device=device, dtype=dtype)

    def put_bn7(key, name):   # BatchNorm1d(7) over (B,7,L): one scale/shift per gem column, applied after the matmul
        s, t = bn(name)
        W[key + "_s"] = s.view(1, 7, 1).to(device=device, dtype=dtype)
        W[key + "_t"] = t.view(1, 7, 1).to(device=device, dtype=dtype)

    put("l1", *lin("dense2d_1.0")); put_bn7("bn1", "dense2d_1.1")
    put("l2", *lin("dense2d_1.3"))
    put("g1", *lin("partialgpool_1.dense_part.0")); put_bn7("bng1", "partialgpool_1.dense_part.1")
    put("l3", *lin("dense2d_3.0"))
    put("l4", *lin("dense1d_4.0"))
    for key, ln, bnn in (("g4", "partialgpool_4.dense_part.0", "partialgpool_4.dense_part.1"),
                         ("l5a", "dense1d_5.0", "dense1d_5.1"), ("g5", "partialgpool_5.dense_part.0", "partialgpool_5.dense_part.1")):
        w, b = lin(ln); s, t = bn(bnn)          # BatchNorm1d(1): a scalar affine, folded into the matrix
        put(key, w * s, b * s + t)
    put("l5b", *lin("dense1d_5.3"))
    for head in ("PI", "V"):
        put(f"{head}0", *lin(f"output_layers_{head}.0")); put(f"{head}1", *lin(f"output_layers_{head}.1"))
    return W


def forward_folded(W, states, valids):
    """states int8[B,R,7], valids uint8/bool[B,406] -> (pi float32[B,406], v float32[B,n]); pure torch, any device"""
    B, R = states.shape[0], states.shape[1]
    dtype = W["l1_w"].dtype
    relu = torch.relu

    def mm(x, key):
        return torch.addmm(W[key + "_b"], x, W[key + "_w"])

    def pool_dense(h, groups, items, key):   # DenseAndPartialGPool (SplendorNNet.py:6-29) on (rows, 128)
        g = h[:, :groups * items].reshape(-1, groups, items)
        return g.amax(-1), g.mean(-1), mm(h[:, groups * items:].contiguous(), key)

    x = states.to(dtype).transpose(1, 2).reshape(B * 7, R)                       # (B,7,R): one row per gem column
    h = mm(x, "l1").view(B, 7, 128)
    h = relu(h * W["bn1_s"] + W["bn1_t"]).view(B * 7, 128)
    h = relu(mm(h, "l2"))
    mx, av, d = pool_dense(h, 4, 8, "g1")
    d = relu(d.view(B, 7, 120) * W["bng1_s"] + W["bng1_t"]).view(B * 7, 120)
    h = torch.cat([mx, av, d], 1)
    h = relu(mm(h, "l3")).view(B, 7, 128)
    first = h[:, :5, :64]                                                          # FlattenAndPartialGPool(64, 5) :32-54
    h = torch.cat([first.amax(1), first.mean(1), h[:, 5:, :64].reshape(B, 128), h[:, :, 64:].reshape(B, 448)], 1)
    h = relu(mm(h, "l4"))
    mx, av, d = pool_dense(h, 4, 4, "g4")
    h = torch.cat([mx, av, relu(d)], 1)
    h = relu(mm(h, "l5a"))
    h = relu(mm(h, "l5b"))
    mx, av, d = pool_dense(h, 4, 4, "g5")
    h = torch.cat([mx, av, relu(d)], 1)
    logits = mm(mm(h, "PI0"), "PI1").float()
    v = torch.tanh(mm(mm(h, "V0"), "V1").float())
    logits = torch.where(valids.bool(), logits, torch.full_like(logits, -1e8))
    return torch.softmax(logits, dim=1), v


PACK_ORDER = (
    [f"dense2d_1.0.{k}" for k in ("weight", "bias")] + [f"dense2d_1.1.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    + [f"dense2d_1.3.{k}" for k in ("weight", "bias")]
    + [f"partialgpool_1.dense_part.0.{k}" for k in ("weight", "bias")]
    + [f"partialgpool_1.dense_part.1.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    + [f"dense2d_3.0.{k}" for k in ("weight", "bias")] + [f"dense1d_4.0.{k}" for k in ("weight", "bias")]
    + [f"partialgpool_4.dense_part.0.{k}" for k in ("weight", "bias")]
    + [f"partialgpool_4.dense_part.1.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    + [f"dense1d_5.0.{k}" for k in ("weight", "bias")] + [f"dense1d_5.1.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    + [f"dense1d_5.3.{k}" for k in ("weight", "bias")]
    + [f"partialgpool_5.dense_part.0.{k}" for k in ("weight", "bias")]
    + [f"partialgpool_5.dense_part.1.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    + [f"output_layers_PI.0.{k}" for k in ("weight", "bias")] + [f"output_layers_PI.1.{k}" for k in ("weight", "bias")]
    + [f"output_layers_V.0.{k}" for k in ("weight", "bias")] + [f"output_layers_V.1.{k}" for k in ("weight", "bias")]
)   # the 46 tensors spl_nnet_pack takes, in order (include/splendor_b200.h)


def pack_blob(n_players, sd):
    """state_dict -> packed weight blob (host uint8 tensor) through the library's own packer (pure host code)"""
    import ctypes as C
    from . import _native as nat
    lib = nat.lib()
    arrs = [np.ascontiguousarray(sd[k].detach().cpu().numpy(), dtype=np.float32) for k in PACK_ORDER]
    ptrs = (C.POINTER(C.c_float) * len(arrs))(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
    nbytes = lib.spl_nnet_blob_bytes(int(n_players))
    blob = np.zeros(nbytes, dtype=np.uint8)
    nat.check(lib.spl_nnet_pack(int(n_players), ptrs, C.c_void_p(blob.ctypes.data), nbytes))
    return torch.from_numpy(blob)


def load_checkpoint_file(path, allow_pickle=False):
    """a checkpoint of GenericNNetWrapper.save_checkpoint (:185-198): {'state_dict', 'full_model', **training args}. Only the
    weights are needed, so the file is read with weights_only=True first; the reference also pickles the whole module
    ('full_model'), which that mode refuses - then only the tensors of 'state_dict' are pulled out of the archive with an
    unpickler that builds nothing but tensors and plain containers (unknown classes become inert placeholders).
    allow_pickle=True: plain torch.load (executes whatever the file's pickle says; needs the reference package importable)."""
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        if allow_pickle:
            return torch.load(path, map_location="cpu", weights_only=False)
    import pickle

    class _Inert:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, state):
            pass

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith(("torch", "collections", "numpy")) or module == "builtins":
                return super().find_class(module, name)
            return _Inert

    class _Pickle:
        Unpickler = _Unpickler
        load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
        __name__ = "pickle"
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Pickle)


def state_dict_from_npz(path):
    """weights stored as an .npz with keys 'sd/<name>' plus the list 'sd_keys' (how tests/golden keeps a checkpoint) -> dict of tensors"""
    z = np.load(path)
    return {str(k): torch.from_numpy(z["sd/" + str(k)]) for k in z["sd_keys"]}


class FusedSplendorNNet:
    """The same evaluator as SplendorNNetB200 in ONE kernel launch (csrc/spl_nnet.cu): bf16 tensor-core products with
    fp32 accumulation, activations resident in shared memory. Output buffers are static per batch size (CUDA-graph safe)."""

    def __init__(self, n_players, state_dict=None, seed=0, device=0):
        import ctypes as C
        from . import _native as nat
        if not torch.cuda.is_available():
            raise RuntimeError("FusedSplendorNNet needs a CUDA device; there is no CPU fallback")
        self.n, self.R = int(n_players), _rows(int(n_players))
        self.device = torch.device("cuda", device)
        self._lib, self._nat, self._C = nat.lib(), nat, C
        h = C.c_void_p()
        nat.check(self._lib.spl_ctx_create(self.n, 10, nat.RULES_DEFAULT, device, C.byref(h)))
        self._ctx = h
        self._out = {}
        self.load_state_dict(state_dict if state_dict is not None else random_state_dict(self.n, seed))
        self.launch_estimate = 1

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                self._lib.spl_ctx_destroy(self._ctx); self._ctx = None
        except Exception:
            pass

    def load_state_dict(self, sd):
        blob = pack_blob(self.n, sd)
        self.blob = torch.empty(blob.numel() + 16, dtype=torch.uint8, device=self.device)
        off = (-self.blob.data_ptr()) % 16
        self._blob_view = self.blob[off:off + blob.numel()]
        self._blob_view.copy_(blob)

    def out_buffers(self, B):
        out = self._out.get(B)
        if out is None:
            out = (torch.zeros((B, NUM_ACTIONS), dtype=torch.float32, device=self.device),
                   torch.zeros((B, self.n), dtype=torch.float32, device=self.device))
            self._out[B] = out
        return out

    @property
    def blob_ptr(self):
        return self._blob_view.data_ptr()

    def __call__(self, states, valids):
        B = states.shape[0]
        out = self.out_buffers(B)
        C = self._C
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._nat.check(self._lib.spl_nnet_forward(self._ctx, C.c_void_p(self._blob_view.data_ptr()), C.c_void_p(states.data_ptr()),
                                                   C.c_void_p(valids.data_ptr()), B, C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), st))
        return out

    def predict(self, board, valid_actions):
        st = torch.from_numpy(np.ascontiguousarray(board, dtype=np.int8)).view(1, self.R, 7).to(self.device)
        va = torch.from_numpy(np.ascontiguousarray(valid_actions).astype(np.uint8)).view(1, NUM_ACTIONS).to(self.device)
        pi, v = self(st, va)
        return pi[0].cpu().numpy(), v[0].cpu().numpy()


class SplendorNNetB200:
    """`predict`-compatible evaluator (NeuralNet.py:33-46) plus the batched device call the tree arena uses."""

    def __init__(self, n_players, state_dict=None, seed=0, device=0, dtype=torch.float32, max_batch=None):
        if not torch.cuda.is_available():
            raise RuntimeError("SplendorNNetB200 needs a CUDA device; there is no CPU fallback")
        self.n = int(n_players)
        self.R = _rows(self.n)
        self.device = torch.device("cuda", device)
        self.dtype = dtype
        self.load_state_dict(state_dict if state_dict is not None else random_state_dict(self.n, seed))
        self.launch_estimate = 40   # torch kernels per forward (bench's launch accounting)

    def load_state_dict(self, sd):
        shapes = state_dict_shapes(self.n)
        for k, shape in shapes.items():
            if k.startswith("output_layers_SDIFF") or k.endswith("num_batches_tracked") or k == "lowvalue":
                continue
            if k not in sd or tuple(sd[k].shape) != tuple(shape):
                raise ValueError(f"state_dict entry {k}: expected shape {shape}, got {tuple(sd[k].shape) if k in sd else None}")
        self.W = fold(sd, self.device, self.dtype)

    def load_checkpoint(self, folder, filename, allow_pickle=False):
        """reads the 'state_dict' entry of a checkpoint written by GenericNNetWrapper.save_checkpoint (:185-198)"""
        import os
        ck = load_checkpoint_file(os.path.join(folder, filename), allow_pickle=allow_pickle)
        self.load_state_dict(ck["state_dict"] if "state_dict" in ck else ck)
        return ck

    @torch.no_grad()
    def __call__(self, states, valids):
        return forward_folded(self.W, states, valids)

    @torch.no_grad()
    def predict(self, board, valid_actions):
        """GenericNNetWrapper.predict (:141-168): one board -> (float32[406] probabilities, float32[n])"""
        st = torch.from_numpy(np.ascontiguousarray(board, dtype=np.int8)).view(1, self.R, 7).to(self.device)
        va = torch.from_numpy(np.ascontiguousarray(valid_actions).astype(np.uint8)).view(1, NUM_ACTIONS).to(self.device)
        pi, v = forward_folded(self.W, st, va)
        return pi[0].cpu().numpy(), v[0].cpu().numpy()
