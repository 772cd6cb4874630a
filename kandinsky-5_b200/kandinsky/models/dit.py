"""Host-side mirror of the reference's `kandinsky/models/dit.py`: same class name, constructor arguments,
`forward` signature and state-dict contract, but the forward is ONE call into libk5 (k5_dit_forward) instead
of ~40 nn.Modules.  The module owns no torch parameters: `load_state_dict` hands every tensor to the engine,
which repacks it (fused QKV, bf16 GEMM operands, fp32 modulation / norms) into its own HBM storage."""
import ctypes
import math
from ctypes import c_float, c_int32, c_int64, c_void_p

import torch
from torch import nn

from .. import _lib
from .._lib import K5Config, K5Sparse, check, lib, ptr, stream_ptr

_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def get_freqs(dim, max_period=10000.0):
    """Same arithmetic as the reference buffer (models/utils.py:21-28): fp32 exp of fp32 arguments, on the CPU."""
    return torch.exp(-math.log(max_period) * torch.arange(start=0, end=dim, dtype=torch.float32) / dim)


def state_dict_shapes(cfg):
    """Checkpoint key -> shape contract of the reference model (dit.py:82-127)."""
    D, F, Td = cfg["model_dim"], cfg["ff_dim"], cfg["time_dim"]
    hd = sum(cfg["axes_dims"])
    cin = 2 * cfg["in_visual_dim"] + 1 if cfg["visual_cond"] else cfg["in_visual_dim"]
    pp = math.prod(cfg["patch_size"])
    s = {}

    def lin(name, o, i, bias=True):
        s[name + ".weight"] = (o, i)
        if bias:
            s[name + ".bias"] = (o,)

    def attn(p):
        for n in ("to_query", "to_key", "to_value", "out_layer"):
            lin(p + n, D, D)
        s[p + "query_norm.weight"] = (hd,)
        s[p + "key_norm.weight"] = (hd,)

    lin("time_embeddings.in_layer", Td, D)
    lin("time_embeddings.out_layer", Td, Td)
    lin("text_embeddings.in_layer", D, cfg["in_text_dim"])
    s["text_embeddings.norm.weight"] = (D,)
    s["text_embeddings.norm.bias"] = (D,)
    lin("pooled_text_embeddings.in_layer", Td, cfg["in_text_dim2"])
    s["pooled_text_embeddings.norm.weight"] = (Td,)
    s["pooled_text_embeddings.norm.bias"] = (Td,)
    lin("visual_embeddings.in_layer", D, pp * cin)
    for i in range(cfg["num_text_blocks"]):
        p = f"text_transformer_blocks.{i}."
        lin(p + "text_modulation.out_layer", 6 * D, Td)
        attn(p + "self_attention.")
        lin(p + "feed_forward.in_layer", F, D, bias=False)
        lin(p + "feed_forward.out_layer", D, F, bias=False)
    for i in range(cfg["num_visual_blocks"]):
        p = f"visual_transformer_blocks.{i}."
        lin(p + "visual_modulation.out_layer", 9 * D, Td)
        attn(p + "self_attention.")
        attn(p + "cross_attention.")
        lin(p + "feed_forward.in_layer", F, D, bias=False)
        lin(p + "feed_forward.out_layer", D, F, bias=False)
    lin("out_layer.modulation.out_layer", 2 * D, Td)
    lin("out_layer.out_layer", pp * cfg["out_visual_dim"], D)
    return s


class DiffusionTransformer3D(nn.Module):
    """Drop-in for `kandinsky.models.dit.DiffusionTransformer3D` (dit.py:82-181) backed by the CUDA engine."""

    def __init__(self, in_visual_dim=4, in_text_dim=3584, in_text_dim2=768, time_dim=512, out_visual_dim=4,
                 patch_size=(1, 2, 2), model_dim=2048, ff_dim=5120, num_text_blocks=2, num_visual_blocks=32,
                 axes_dims=(16, 24, 24), visual_cond=False, max_tokens=47616, max_text_tokens=512):
        super().__init__()
        self.cfg = dict(in_visual_dim=in_visual_dim, in_text_dim=in_text_dim, in_text_dim2=in_text_dim2,
                        time_dim=time_dim, out_visual_dim=out_visual_dim, patch_size=tuple(patch_size),
                        model_dim=model_dim, ff_dim=ff_dim, num_text_blocks=num_text_blocks,
                        num_visual_blocks=num_visual_blocks, axes_dims=tuple(axes_dims), visual_cond=bool(visual_cond))
        self.in_visual_dim = in_visual_dim
        self.model_dim = model_dim
        self.patch_size = tuple(patch_size)
        self.visual_cond = bool(visual_cond)
        self.max_tokens = int(max_tokens)
        self.max_text_tokens = int(max_text_tokens)
        self._engine = None
        self._device = None
        self._grid_key = None
        self._pending = None          # CPU state dict kept until .to(cuda) creates the engine
        self.dist_rank, self.dist_world = 0, 1
        self._magcache = None         # magcache_utils.MagCacheState when get_T2V_pipeline(magcache=True)

    # ---- engine lifetime -------------------------------------------------------------------------------
    def _create_engine(self, device):
        c = K5Config()
        cfg = self.cfg
        c.in_visual_dim, c.out_visual_dim, c.time_dim = cfg["in_visual_dim"], cfg["out_visual_dim"], cfg["time_dim"]
        c.patch_size = (c_int32 * 3)(*cfg["patch_size"])
        c.model_dim, c.ff_dim = cfg["model_dim"], cfg["ff_dim"]
        c.num_text_blocks, c.num_visual_blocks = cfg["num_text_blocks"], cfg["num_visual_blocks"]
        c.axes_dims = (c_int32 * 3)(*cfg["axes_dims"])
        c.visual_cond = int(cfg["visual_cond"])
        c.in_text_dim, c.in_text_dim2 = cfg["in_text_dim"], cfg["in_text_dim2"]
        c.max_tokens, c.max_text_tokens = self.max_tokens, self.max_text_tokens
        handle = c_void_p()
        with torch.cuda.device(device):
            check(lib().k5_engine_create(ctypes.byref(c), ctypes.byref(handle)))
        self._engine, self._device = handle, torch.device(device)

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng is not None and _lib._lib is not None:
            _lib._lib.k5_engine_destroy(eng)
            self._engine = None

    def _ref_buffers(self):
        """The reference's non-persistent buffers (nn.py:49-51,107,129), computed the way it computes them."""
        cfg = self.cfg
        hd = sum(cfg["axes_dims"])
        out = {"time_embeddings.freqs": get_freqs(cfg["model_dim"] // 2),
               "text_rope_embeddings.args": torch.outer(torch.arange(1024, dtype=torch.float32), get_freqs(hd // 2))}
        for i, ad in enumerate(cfg["axes_dims"]):
            out[f"visual_rope_embeddings.args_{i}"] = torch.outer(torch.arange(128, dtype=torch.float32),
                                                                  get_freqs(ad // 2))
        return out

    def _load_one(self, key, t):
        t = t.detach()
        if t.dtype not in _DTYPE_CODE:
            t = t.float()
        t = t.contiguous()
        shape = (c_int64 * t.dim())(*t.shape)
        check(lib().k5_engine_load_tensor(self._engine, key.encode(), c_void_p(t.data_ptr()), _DTYPE_CODE[t.dtype], shape,
                                          t.dim()))

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Same contract as the reference's `dit.load_state_dict(state_dict, assign=True)` (utils.py:115-116):
        exactly the 814 keys (for the Lite config), any of fp32 / bf16 / fp16."""
        want = state_dict_shapes(self.cfg)
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                               f"unexpected keys {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
        for k, shp in want.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model {tuple(shp)}")
        if self._engine is None:
            self._pending = {k: v for k, v in state_dict.items() if k in want}
        else:
            self._upload({k: v for k, v in state_dict.items() if k in want})
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _upload(self, sd):
        with torch.cuda.device(self._device):
            for k, v in sd.items():
                self._load_one(k, v)
            for k, v in self._ref_buffers().items():
                self._load_one(k, v)
            check(lib().k5_engine_finalize(self._engine))

    def to(self, device=None, *args, **kwargs):
        if device is not None and torch.device(device).type == "cuda":
            dev = torch.device(device)
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            if self._engine is None:
                self._create_engine(dev)
                if self._pending is not None:
                    self._upload(self._pending)
                    self._pending = None
            elif dev != self._device:
                raise RuntimeError("the engine is bound to one GPU; create a new model for another device")
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        return self

    # ---- temporal shard (include/k5.h k5_dist_*; replaces models/parallelize.py of the reference) -----------
    def dist_export(self):
        """Opaque handle (bytes) of this rank's K|V exchange buffer; exchange them between ranks, then dist_init."""
        if self._engine is None:
            raise RuntimeError("dist_export: the model must be on its CUDA device first")
        buf = ctypes.create_string_buffer(_lib.DIST_HANDLE_BYTES)
        with torch.cuda.device(self._device):
            check(lib().k5_dist_export(self._engine, buf))
        return buf.raw

    def dist_init(self, rank, world, handles):
        """handles: list of `world` blobs from dist_export, in rank order."""
        if len(handles) != world or any(len(h) != _lib.DIST_HANDLE_BYTES for h in handles):
            raise ValueError("dist_init: need one exported handle per rank")
        blob = ctypes.create_string_buffer(b"".join(handles), world * _lib.DIST_HANDLE_BYTES)
        with torch.cuda.device(self._device):
            check(lib().k5_dist_init(self._engine, int(rank), int(world), blob))
        self.dist_rank, self.dist_world = int(rank), int(world)
        self._grid_key = None

    def local_frames(self):
        """(first_frame, num_frames) of the latent this rank owns for the current grid."""
        f0, n = ctypes.c_int(0), ctypes.c_int(0)
        check(lib().k5_dist_local_frames(self._engine, ctypes.byref(f0), ctypes.byref(n)))
        return f0.value, n.value

    def dist_mode(self):
        """0 = single GPU, 1 = K|V scatter from the QKV epilogue + flag barrier, 2 = overlapped all-gather (k5_dist_mode)."""
        return int(lib().k5_dist_mode(self._engine)) if self._engine is not None else 0

    # ---- forward ---------------------------------------------------------------------------------------
    def set_grid(self, shape, visual_rope_pos, scale_factor, fractal):
        T, H, W = shape
        if len(visual_rope_pos) != 3 or [len(p) for p in visual_rope_pos] != [T, H // 2, W // 2]:
            raise ValueError(f"visual_rope_pos must hold {T}, {H // 2}, {W // 2} positions (dit.py:161), got "
                             f"{[len(p) for p in visual_rope_pos]}")
        key = (T, H, W, tuple(float(s) for s in scale_factor), bool(fractal),
               tuple(tuple(int(v) for v in p.tolist()) for p in visual_rope_pos))
        if key == self._grid_key:
            return
        pos = [(c_int32 * len(p))(*[int(v) for v in p.tolist()]) for p in visual_rope_pos]
        sf = (c_float * 3)(*[float(s) for s in scale_factor])
        check(lib().k5_engine_set_grid(self._engine, T, H, W, pos[0], pos[1], pos[2], sf, 1 if fractal else 0))
        self._grid_key = key

    @staticmethod
    def _sparse_struct(sparse_params):
        if sparse_params is None:
            return None
        sp = K5Sparse()
        sp.P = float(sparse_params["P"])
        sp.wT, sp.wH, sp.wW = int(sparse_params["wT"]), int(sparse_params["wH"]), int(sparse_params["wW"])
        # nablaT_v2 (models/utils.py:152) ORs the STA mask unconditionally - the reference never reads `add_sta` - so the
        # engine does the same; the key is accepted for compatibility
        sp.add_sta = 1
        return sp

    @torch.no_grad()
    def forward(self, x, text_embed, pooled_text_embed, time, visual_rope_pos, text_rope_pos,
                scale_factor=(1.0, 1.0, 1.0), sparse_params=None):
        """Same arguments as the reference forward (dit.py:156-166).  x [T,H,W,C] fp32 -> [T,H,W,out] bf16."""
        if self._engine is None:
            raise RuntimeError("model is not on a CUDA device / no weights loaded: call load_state_dict() and .to('cuda')")
        if x.dim() != 4:
            raise ValueError("x must be [T, H, W, C]")
        T, H, W, C = x.shape
        fractal = bool(sparse_params["to_fractal"]) if sparse_params is not None else False
        with torch.cuda.device(self._device):
            self.set_grid((T, H, W), visual_rope_pos, scale_factor, fractal)
            x = x.to(self._device, torch.float32).contiguous()
            text = text_embed.to(self._device, torch.bfloat16).contiguous()
            pooled = pooled_text_embed.to(self._device, torch.bfloat16).contiguous().view(-1)
            L = text.shape[0]
            tpos = [int(v) for v in text_rope_pos.tolist()]
            tpos_arr = None if tpos == list(range(L)) else (c_int32 * L)(*tpos)
            # on a temporal shard the engine writes this rank's frames only (the slabs are gathered once, after the last
            # step: parallelize.gather_frames); the other frames read as zeros, never as uninitialised memory
            alloc = torch.zeros if self.dist_world > 1 else torch.empty
            out = alloc(T, H, W, self.cfg["out_visual_dim"], device=self._device, dtype=torch.bfloat16)
            sp = self._sparse_struct(sparse_params)
            t = float(time.reshape(-1)[0].item()) if torch.is_tensor(time) else float(time)
            spp = ctypes.byref(sp) if sp is not None else None
            if self._magcache is None:
                check(lib().k5_dit_forward(self._engine, ptr(x), C, ptr(text), L, tpos_arr, ptr(pooled), t, spp, ptr(out),
                                           stream_ptr()))
            else:
                slot, skip = self._magcache.next()
                check(lib().k5_dit_forward_magcache(self._engine, ptr(x), C, ptr(text), L, tpos_arr, ptr(pooled), t, spp,
                                                    ptr(out), slot, 1 if skip else 0, stream_ptr()))
                self.last_magcache_decision = (slot, skip)
        return out

    __call__ = forward

    def last_sparse_density(self):
        return float(lib().k5_last_sparse_density(self._engine))


def get_dit(conf):
    """dit.py:184-186."""
    return DiffusionTransformer3D(**conf)
