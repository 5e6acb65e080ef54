"""Host side of the B200 engine: torch tensors in, raw device pointers across the C ABI.

PyTorch is used for device memory, streams and (in parallel.py) torch.distributed only; all
arithmetic of the hot path runs in libb200dit.so.  `DitEngine.forward` has the exact argument
meaning of the reference `WanModel.forward` (seaweed_apt/wan/modules/model.py:502-563),
`VaeEngine.decode` of `WanVAE.decode` (seaweed_apt/wan/modules/vae.py:657-663).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import B200Error, DitConfig, check, int_array, lib, ptr_array

_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _load_state(load_fn, handle, sd, accept):
    """One load call per state_dict entry.  Device-resident sources are converted asynchronously (the engine
    synchronises once in *_finalize), so every tensor handed over is returned and must stay referenced until then."""
    keep = []
    for name, tensor in sd.items():
        if not accept(name):
            continue
        t = tensor.detach()
        if t.dtype not in _DT:
            t = t.float()
        t = t.contiguous()
        keep.append(t)
        shape = (C.c_int64 * max(t.dim(), 1))(*(list(t.shape) or [1]))
        check(load_fn(handle, name.encode(), C.c_void_p(t.data_ptr()), _DT[t.dtype], max(t.dim(), 1), shape))
    return keep


class DitEngine:
    """B200 replacement for the arithmetic of one `WanModel` instance."""

    def __init__(self, dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16, text_dim=4096,
                 text_len=512, freq_dim=256, i2v=False, eps=1e-6, device=None):
        if not torch.cuda.is_available():
            raise B200Error("no CUDA device: the B200 engine has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.cfg = dict(dim=dim, ffn_dim=ffn_dim, num_heads=num_heads, num_layers=num_layers, in_dim=in_dim,
                        out_dim=out_dim, text_dim=text_dim, text_len=text_len, freq_dim=freq_dim, i2v=int(bool(i2v)),
                        eps=eps)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().b200dit_create(C.byref(DitConfig(**self.cfg)), C.byref(self._h)))
        self._tap = None
        self._ctx_key, self._ctx_refs, self._ctx_tok = None, None, 0
        self.cache_context = True

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_state_dict(cls, sd, num_heads, text_len=512, eps=1e-6, device=None):
        """Infers the architecture from reference state_dict shapes (key names: model.py:463-498)."""
        dim = sd["patch_embedding.weight"].shape[0]
        layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
        eng = cls(dim=dim, ffn_dim=sd["blocks.0.ffn.0.weight"].shape[0], num_heads=num_heads, num_layers=layers,
                  in_dim=sd["patch_embedding.weight"].shape[1], out_dim=sd["head.head.weight"].shape[0] // 4,
                  text_dim=sd["text_embedding.0.weight"].shape[1], text_len=text_len,
                  freq_dim=sd["time_embedding.0.weight"].shape[1], i2v="img_emb.proj.0.weight" in sd, eps=eps,
                  device=device)
        eng.load_state_dict(sd)
        return eng

    @staticmethod
    def config_from_module(model):
        """Architecture of a reference `WanModel` instance from the attributes its constructor sets
        (model.py:445-460).  Pure Python: runs without a GPU (tests/test_cpu_boundary.py drives it with the real
        class)."""
        return dict(dim=model.dim, ffn_dim=model.ffn_dim, num_heads=model.num_heads, num_layers=model.num_layers,
                    in_dim=model.in_dim, out_dim=model.out_dim, text_dim=model.text_dim, text_len=model.text_len,
                    freq_dim=model.freq_dim, i2v=(model.model_type == "i2v"), eps=model.eps)

    @staticmethod
    def expected_weight_names(**cfg):
        """{state_dict key: element count} an engine of this architecture loads (b200dit_weight_names; host only)."""
        full = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16, text_dim=4096,
                    text_len=512, freq_dim=256, i2v=0, eps=1e-6)
        full.update(cfg)
        full["i2v"] = int(bool(full["i2v"]))
        c = DitConfig(**full)
        need = lib().b200dit_weight_names(C.byref(c), None, 0)
        if need < 0:
            raise B200Error(lib().b200_last_error().decode("utf-8", "replace"))
        buf = C.create_string_buffer(int(need))
        lib().b200dit_weight_names(C.byref(c), buf, need)
        out = {}
        for line in buf.value.decode().splitlines():
            name, n = line.rsplit(" ", 1)
            out[name] = int(n)
        return out

    @classmethod
    def from_module(cls, model, device=None):
        """Builds an engine from a reference `WanModel` instance (attributes set at model.py:445-460)."""
        eng = cls(**cls.config_from_module(model), device=device)
        eng.load_state_dict(model.state_dict())
        return eng

    def load_state_dict(self, sd):
        self._ctx_key = None              # cached cross-attention K / V belong to the old weights
        with torch.cuda.device(self.device):
            keep = _load_state(lib().b200dit_load_weight, self._h, sd, lambda n: n != "freqs")
            check(lib().b200dit_finalize(self._h))
            del keep

    def set_pad_to_seq_len(self, enabled):
        """Carry the reference's seq_len - L zero-padded rows of every item through the blocks (model.py:522), so that
        taps hold the [n_items * seq_len, dim] block outputs the APT discriminator reads (b200dit_set_pad_to_seq_len)."""
        check(lib().b200dit_set_pad_to_seq_len(self._h, int(bool(enabled))))

    def set_graphs(self, enabled):
        check(lib().b200dit_set_graphs(self._h, int(bool(enabled))))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().b200dit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def _prep_items(self, x, context, clip_fea, y):
        xs = [u for u in x] if not isinstance(x, (list, tuple)) else list(x)
        xs = [u.to(self.device, torch.float32, non_blocking=True).contiguous() for u in xs]
        ys = None
        if y is not None:
            ys = [v.to(self.device, torch.float32, non_blocking=True).contiguous() for v in y]
        ctx = []
        for c in context:
            c = c if c.dtype in _DT else c.float()
            ctx.append(c.to(self.device, non_blocking=True).contiguous())
        dts = {c.dtype for c in ctx}
        if len(dts) > 1:
            ctx = [c.float() for c in ctx]
        clips = None
        if clip_fea is not None:
            clips = [c.to(self.device, torch.float32, non_blocking=True).contiguous() for c in clip_fea]
        return xs, ys, ctx, clips

    def _t_tensor(self, t, n):
        if not torch.is_tensor(t):
            t = torch.tensor(t, dtype=torch.float32)
        t = t.reshape(-1).to(self.device, torch.float32, non_blocking=True)
        if t.numel() == 1 and n > 1:
            t = t.expand(n)
        assert t.numel() == n, f"t has {t.numel()} entries for {n} items"
        return t.contiguous()

    def _context_hint(self, n_calls, *tensor_lists):
        """Tells the engine when this call carries the very same context tensors as the previous one
        (b200dit_context_hint): same objects, same storage, no in-place update since (`_version`).  The
        tensors are kept referenced so their storage cannot be recycled for different contents."""
        if not self.cache_context or n_calls != 1:
            return
        flat = [c for lst in tensor_lists if lst is not None for c in lst]
        key = tuple((c.data_ptr(), c._version, tuple(c.shape), c.dtype, str(c.device)) for c in flat)
        if key != self._ctx_key:
            self._ctx_key, self._ctx_refs = key, flat
            self._ctx_tok += 1
        check(lib().b200dit_context_hint(self._h, self._ctx_tok))

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None):
        """WanModel.forward (model.py:502): list (or batched tensor) of [C,F,H,W] -> list of fp32 [16,F,H,W]."""
        xs, ys, ctx, clips = self._prep_items(x, context, clip_fea, y)
        n = len(xs)
        tt = self._t_tensor(t, n)
        outs = [None] * n
        groups = {}
        for i, u in enumerate(xs):
            groups.setdefault(tuple(u.shape[1:]), []).append(i)
        # items are co-batched per latent grid; a token count that is not a multiple of 8 cannot be co-batched
        # (every item must start at a 16-byte aligned column of the transposed V) and runs one item per call
        chunk = lambda F, H, W: _lib.MAX_ITEMS if (F * (H // 2) * (W // 2)) % 8 == 0 else 1
        n_calls = sum((len(idx) + chunk(*g) - 1) // chunk(*g) for g, idx in groups.items())
        with torch.cuda.device(self.device):
            for (F, H, W), idx in groups.items():
                for s in range(0, len(idx), chunk(F, H, W)):
                    part = idx[s:s + chunk(F, H, W)]
                    self._context_hint(n_calls, context, clip_fea)
                    o = [torch.empty((self.cfg["out_dim"], F, H, W), dtype=torch.float32, device=self.device)
                         for _ in part]
                    t_part = tt[part].contiguous() if len(part) != n else tt
                    ych = ys[part[0]].shape[0] if ys is not None else 0
                    check(lib().b200dit_forward(
                        self._h, len(part), ptr_array([xs[i].data_ptr() for i in part]),
                        ptr_array([ys[i].data_ptr() for i in part]) if ys is not None else None, ych,
                        C.c_void_p(t_part.data_ptr()), ptr_array([ctx[i].data_ptr() for i in part]),
                        int_array([ctx[i].shape[0] for i in part]), _DT[ctx[part[0]].dtype],
                        ptr_array([clips[i].data_ptr() for i in part]) if clips is not None else None,
                        F, H, W, int(seq_len) if seq_len is not None else 0,
                        ptr_array([q.data_ptr() for q in o]), _stream_ptr()))
                    for i, q in zip(part, o):
                        outs[i] = q
        return outs

    __call__ = forward

    def forward_cfg(self, x, t, context, context_null, seq_len, guide_scale, clip_fea=None, y=None):
        """cond + uncond forwards and `uncond + s (cond - uncond)` (text2video.py:238-244) in one call."""
        xs, ys, ctx, clips = self._prep_items(x, context, clip_fea, y)
        _, _, ctx_n, _ = self._prep_items([], context_null, None, None)
        n = len(xs)
        if len(ctx_n) == 1 and n > 1:
            ctx_n = ctx_n * n
        if ctx and ctx_n and ctx[0].dtype != ctx_n[0].dtype:
            ctx, ctx_n = [c.float() for c in ctx], [c.float() for c in ctx_n]
        tt = self._t_tensor(t, n)
        outs = [None] * n
        groups = {}
        for i, u in enumerate(xs):
            groups.setdefault(tuple(u.shape[1:]), []).append(i)
        if any((F * (H // 2) * (W // 2)) % 8 != 0 for (F, H, W) in groups):
            # token counts that cannot be co-batched: two plain forwards and the combine as one fused launch
            from .solvers import _lincomb
            c = self.forward(xs, tt, ctx, seq_len, clip_fea=clips, y=ys)
            u = self.forward(xs, tt, ctx_n, seq_len, clip_fea=clips, y=ys)
            g = float(guide_scale)
            return [_lincomb([ci, ui], [[g, 1.0 - g]], ci)[0] for ci, ui in zip(c, u)]
        half = _lib.MAX_ITEMS // 2
        n_calls = sum((len(idx) + half - 1) // half for idx in groups.values())
        with torch.cuda.device(self.device):
            for (F, H, W), idx in groups.items():
                for s in range(0, len(idx), half):
                    part = idx[s:s + half]
                    self._context_hint(n_calls, context, context_null, clip_fea)
                    o = [torch.empty((self.cfg["out_dim"], F, H, W), dtype=torch.float32, device=self.device)
                         for _ in part]
                    t_part = tt[part].contiguous() if len(part) != n else tt
                    ych = ys[part[0]].shape[0] if ys is not None else 0
                    check(lib().b200dit_forward_cfg(
                        self._h, len(part), ptr_array([xs[i].data_ptr() for i in part]),
                        ptr_array([ys[i].data_ptr() for i in part]) if ys is not None else None, ych,
                        C.c_void_p(t_part.data_ptr()),
                        ptr_array([ctx[i].data_ptr() for i in part]), int_array([ctx[i].shape[0] for i in part]),
                        ptr_array([ctx_n[i].data_ptr() for i in part]), int_array([ctx_n[i].shape[0] for i in part]),
                        _DT[ctx[part[0]].dtype],
                        ptr_array([clips[i].data_ptr() for i in part]) if clips is not None else None,
                        F, H, W, int(seq_len) if seq_len is not None else 0, float(guide_scale),
                        ptr_array([q.data_ptr() for q in o]), _stream_ptr()))
                    for i, q in zip(part, o):
                        outs[i] = q
        return outs

    def set_tap(self, block_idx, n_rows=None):
        """Residual stream after block `block_idx` (APT discriminator taps, seaweed_apt/model.py:150-155)."""
        if block_idx is None or block_idx < 0:
            check(lib().b200dit_set_tap(self._h, -1, None, 0))
            self._tap = None
            return None
        self._tap = torch.empty((n_rows, self.cfg["dim"]), dtype=torch.float32, device=self.device)
        check(lib().b200dit_set_tap(self._h, int(block_idx), C.c_void_p(self._tap.data_ptr()), int(n_rows)))
        return self._tap

    def set_taps(self, block_indices, n_rows, buffers=None):
        """Residual stream after several blocks (the APT discriminator hooks three, seaweed_apt/model.py:150-155);
        returns one fp32 [n_rows, dim] tensor per block, refreshed by every following forward.  `buffers`
        re-registers tensors returned by an earlier call instead of allocating new ones."""
        idx = [int(b) for b in block_indices]
        if buffers is not None:
            assert len(buffers) == len(idx) and all(b.shape == (n_rows, self.cfg["dim"]) for b in buffers)
            self._tap = list(buffers)
        else:
            self._tap = [torch.empty((n_rows, self.cfg["dim"]), dtype=torch.float32, device=self.device) for _ in idx]
        check(lib().b200dit_set_taps(self._h, len(idx), int_array(idx), ptr_array([t.data_ptr() for t in self._tap]),
                                    int(n_rows)))
        return self._tap

    # ------------------------------------------------------------------ F1: training step (distilled_trainer.py:268-301)
    def train_forward(self, x, t, context, seq_len):
        """`WanModel.forward` for items of ONE latent grid, keeping what `backward` needs (b200dit_train_forward):
        the residual stream at every block boundary, as the reference's per-block checkpointing does
        (model.py:544-548).  t2v only (no `y` / `clip_fea`)."""
        xs, _, ctx, _ = self._prep_items(x, context, None, None)
        n = len(xs)
        assert n >= 1 and all(u.shape == xs[0].shape for u in xs), "train_forward takes items of one latent grid"
        if n > 1 and (xs[0].shape[1] * (xs[0].shape[2] // 2) * (xs[0].shape[3] // 2)) % 8 != 0:
            raise B200Error("co-batched training items need a token count that is a multiple of 8; call once per item")
        tt = self._t_tensor(t, n)
        _, F, H, W = xs[0].shape
        outs = [torch.empty((self.cfg["out_dim"], F, H, W), dtype=torch.float32, device=self.device) for _ in xs]
        with torch.cuda.device(self.device):
            check(lib().b200dit_train_forward(
                self._h, n, ptr_array([u.data_ptr() for u in xs]), C.c_void_p(tt.data_ptr()),
                ptr_array([c.data_ptr() for c in ctx]), int_array([c.shape[0] for c in ctx]), _DT[ctx[0].dtype],
                F, H, W, int(seq_len) if seq_len is not None else 0, ptr_array([o.data_ptr() for o in outs]),
                _stream_ptr()))
        self._train_refs = (xs, tt, ctx)                  # inputs stay alive until the backward has run
        self._train_seq = getattr(self, "_train_seq", 0) + 1
        return outs

    def backward(self, douts, loss_scale=None, ffn_grad_blocks=11, want_dx=True):
        """`loss.backward()` through the latest `train_forward` (b200dit_backward).  douts: d loss / d out per item.
        Parameter gradients accumulate inside the engine (`read_grad`, `zero_grad`); returns d loss / d x per item
        (or None).  `ffn_grad_blocks=11` reproduces the reference, whose blocks with block_idx > 10 run their FFN
        under no_grad (model.py:318-325); None differentiates every FFN.  `loss_scale=None` picks a power of two
        that puts max|dout| near 256 (the contractions carry fp16 operands, like the reference under its
        GradScaler, distilled_trainer.py:88,301; the stored gradients are unscaled); one device->host read."""
        ds = [d.to(self.device, torch.float32).contiguous() for d in douts]
        if loss_scale is None:
            amax = float(torch.stack([d.abs().max() for d in ds]).max())
            loss_scale = 1.0 if not (amax > 0.0 and amax < float("inf")) else 2.0 ** int(torch.floor(torch.log2(torch.tensor(256.0 / amax))))
        dxs = [torch.empty_like(u) for u in self._train_refs[0]] if want_dx else None
        with torch.cuda.device(self.device):
            check(lib().b200dit_backward(
                self._h, ptr_array([d.data_ptr() for d in ds]), float(loss_scale),
                -1 if ffn_grad_blocks is None else int(ffn_grad_blocks),
                ptr_array([u.data_ptr() for u in dxs]) if want_dx else None, _stream_ptr()))
        return dxs

    def zero_grad(self):
        with torch.cuda.device(self.device):
            check(lib().b200dit_zero_grad(self._h, _stream_ptr()))

    def grad_buffers(self):
        """The engine's two gradient stores as zero-copy fp32 CUDA tensors (b200dit_grad_buffers): what a
        data-parallel trainer all-reduces between `backward` and its optimizer step (`parallel.all_reduce_gradients`)."""
        p16, p32, n16, n32 = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        with torch.cuda.device(self.device):
            check(lib().b200dit_grad_buffers(self._h, C.byref(p16), C.byref(n16), C.byref(p32), C.byref(n32)))

        class _Raw:                                   # __cuda_array_interface__: borrowed device memory, no copy
            def __init__(s, ptr, n):
                s.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f4", data=(int(ptr), False), version=3)

        return [torch.as_tensor(_Raw(p16.value, n16.value), device=self.device),
                torch.as_tensor(_Raw(p32.value, n32.value), device=self.device)]

    def read_grad(self, name, shape, out=None, accumulate=False):
        """Gradient of the parameter stored under the reference key `name` (fp32), b200dit_read_grad."""
        scale = 1.0
        if out is None:
            out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
            accumulate = False
        assert out.is_contiguous() and out.dtype == torch.float32 and out.device == self.device
        with torch.cuda.device(self.device):
            check(lib().b200dit_read_grad(self._h, name.encode(), C.c_void_p(out.data_ptr()), out.numel(), scale,
                                          int(bool(accumulate)), _stream_ptr()))
        return out

    @property
    def last_flops(self):
        return float(lib().b200dit_last_flops(self._h))

    def nonfinite_rows(self):
        """Overflow guard: residual-stream rows that reached a LayerNorm as inf / NaN since the previous call
        (GEMM / attention operands are fp16, like the reference's blocks under autocast(float16), model.py:540).  0 on a healthy run.  Blocks
        until the current stream drains."""
        n = C.c_uint32(0)
        check(lib().b200dit_nonfinite_rows(self._h, torch.cuda.current_stream(self.device).cuda_stream, C.byref(n)))
        return int(n.value)


class VaeEngine:
    """B200 replacement for `WanVAE.decode` (vae.py:657-663)."""

    def __init__(self, dim=96, z_dim=16, device=None):
        if not torch.cuda.is_available():
            raise B200Error("no CUDA device: the B200 engine has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.dim, self.z_dim = dim, z_dim
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().b200vae_create(dim, z_dim, C.byref(self._h)))

    @staticmethod
    def config_from_state_dict(sd):
        """(dim, z_dim) of a `WanVAE_` from its state_dict shapes (vae.py:388-421,454-472); no GPU needed."""
        return dict(dim=sd["decoder.head.2.weight"].shape[1], z_dim=sd["conv2.weight"].shape[0])

    @classmethod
    def from_state_dict(cls, sd, device=None):
        eng = cls(**cls.config_from_state_dict(sd), device=device)
        eng.load_state_dict(sd)
        return eng

    def load_state_dict(self, sd, encoder=None):
        """`decoder.*` / `conv2.*` always; `encoder.*` / `conv1.*` too when present (encoder=None) or demanded."""
        self.has_encoder = any(k.startswith("encoder.") for k in sd) if encoder is None else bool(encoder)
        prefixes = ("decoder.", "conv2.") + (("encoder.", "conv1.") if self.has_encoder else ())
        with torch.cuda.device(self.device):
            _load_state(lib().b200vae_load_weight, self._h, sd, lambda n: n.startswith(prefixes))
            check(lib().b200vae_finalize(self._h))

    def encode(self, videos):
        """WanVAE.encode (vae.py:641-655): list of [3, T, H, W] in [-1, 1] (T = 1 + 4k) -> list of fp32
        [16, 1 + k, H/8, W/8] posterior means, normalised with the latent mean / std."""
        outs = []
        with torch.cuda.device(self.device):
            for v in videos:
                v = v.to(self.device, torch.float32, non_blocking=True).contiguous()
                _, T, H, W = v.shape
                o = torch.empty((self.z_dim, 1 + (T - 1) // 4, H // 8, W // 8), dtype=torch.float32, device=self.device)
                check(lib().b200vae_encode(self._h, C.c_void_p(v.data_ptr()), T, H, W, C.c_void_p(o.data_ptr()),
                                           _stream_ptr()))
                outs.append(o)
        return outs

    def decode(self, zs):
        """List of [16,T,h,w] latents -> list of fp32 [3, 1+4(T-1), 8h, 8w] in [-1, 1]."""
        outs = []
        with torch.cuda.device(self.device):
            for z in zs:
                z = z.to(self.device, torch.float32, non_blocking=True).contiguous()
                _, T, h, w = z.shape
                o = torch.empty((3, 1 + 4 * (T - 1), 8 * h, 8 * w), dtype=torch.float32, device=self.device)
                check(lib().b200vae_decode(self._h, C.c_void_p(z.data_ptr()), T, h, w, C.c_void_p(o.data_ptr()),
                                           _stream_ptr()))
                outs.append(o)
        return outs

    def decode_pipelined(self, z, chunk_frames=None, group=None):
        """Multi-GPU time-chunked decode of ONE latent [16, T, h, w]: every rank of the default
        process group calls this with the same latent and gets the whole fp32 video [3, 1 + 4 (T - 1), 8h, 8w].
        Rank r decodes chunks r, r + world, ... and hands the per-conv two-frame caches (vae.py:207-217) to the
        next rank through peer memory (b200vae_decode_pipelined); one all_reduce assembles the frames.  `group`: a
        process group of ranks on one node (default: all ranks), e.g. parallel.pair_group()."""
        import torch.distributed as dist
        from . import parallel
        if parallel.world_size() == 1:
            return self.decode([z])[0]
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return self.decode([z])[0]
        z = z.to(self.device, torch.float32).contiguous()
        _, T, h, w = z.shape
        with torch.cuda.device(self.device):
            gkey = id(group) if group is not None else 0
            if getattr(self, "_pipe_key", None) != (h, w, world, gkey):
                handle = (C.c_uint8 * 64)()
                check(lib().b200vae_pipe_prepare(self._h, h, w, handle))
                handles = [None] * world
                dist.all_gather_object(handles, bytes(handle), group=group)
                nxt = (C.c_uint8 * 64).from_buffer_copy(handles[(rank + 1) % world])
                check(lib().b200vae_pipe_connect(self._h, nxt))
                dist.barrier(group=group)                       # every arena is mapped before anyone writes
                self._pipe_key, self._pipe_epoch = (h, w, world, gkey), 0
            self._pipe_epoch += 1
            cf = int(chunk_frames or parallel.pipeline_chunk_frames(T, world))
            out = torch.zeros((3, 1 + 4 * (T - 1), 8 * h, 8 * w), dtype=torch.float32, device=self.device)
            check(lib().b200vae_decode_pipelined(self._h, C.c_void_p(z.data_ptr()), T, h, w, C.c_void_p(out.data_ptr()),
                                                 rank, world, cf, self._pipe_epoch, _stream_ptr()))
            return parallel.sum_disjoint(out, group=group)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().b200vae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                    window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """Drop-in for the reference operator seam `flash_attention` (attention.py:24-38): same signature,
    q [B,Lq,N,128], k/v [B,Lk,N,128]; returns q's dtype.  Operands run in fp16 (model.py:540 autocast)."""
    assert q.is_cuda and q.size(-1) == 128, "B200 flash_attention needs CUDA tensors with head_dim 128"
    assert not causal and q_lens is None and dropout_p == 0. and tuple(window_size) == (-1, -1)
    b, lq, n, _ = q.shape
    lk = k.shape[1]
    out_dtype = q.dtype
    if q_scale is not None:
        q = q * q_scale
    qh, kh, vh = (u.to(torch.float16).contiguous() for u in (q, k, v))
    out = torch.empty_like(qh)
    kl = None
    if k_lens is not None:
        kl = int_array([int(u) for u in (k_lens.tolist() if torch.is_tensor(k_lens) else k_lens)])
    with torch.cuda.device(q.device):
        check(lib().b200_flash_attention(C.c_void_p(qh.data_ptr()), C.c_void_p(kh.data_ptr()), C.c_void_p(vh.data_ptr()),
                                         kl, b, lq, lk, n, float(softmax_scale or 0.0), C.c_void_p(out.data_ptr()),
                                         _stream_ptr()))
    return out.to(out_dtype)


def flash_attention_backward(q, k, v, dout, k_lens=None, softmax_scale=None):
    """Adjoint of `flash_attention` (b200_flash_attention_backward): q, dout [B,Lq,N,128], k / v [B,Lk,N,128] ->
    (dq fp32, dk fp32, dv fp16), the gradients autograd would hand back through attention.py:24-130."""
    assert q.is_cuda and q.size(-1) == 128
    b, lq, n, _ = q.shape
    lk = k.shape[1]
    qh, kh, vh, dh = (u.to(torch.float16).contiguous() for u in (q, k, v, dout))
    dq = torch.empty((b, lq, n, 128), dtype=torch.float32, device=q.device)
    dk = torch.empty((b, lk, n, 128), dtype=torch.float32, device=q.device)
    dv = torch.empty((b, lk, n, 128), dtype=torch.float16, device=q.device)
    kl = None
    if k_lens is not None:
        kl = int_array([int(u) for u in (k_lens.tolist() if torch.is_tensor(k_lens) else k_lens)])
    with torch.cuda.device(q.device):
        check(lib().b200_flash_attention_backward(
            C.c_void_p(qh.data_ptr()), C.c_void_p(kh.data_ptr()), C.c_void_p(vh.data_ptr()), C.c_void_p(dh.data_ptr()), kl,
            b, lq, lk, n, float(softmax_scale or 0.0), C.c_void_p(dq.data_ptr()), C.c_void_p(dk.data_ptr()),
            C.c_void_p(dv.data_ptr()), _stream_ptr()))
    return dq, dk, dv


def linear(a, w, bias=None, epilogue="f16", block_n=0):
    """nn.Linear on the tcgen05 GEMM: a [M,K] fp16, w [N,K] fp16 -> [M,N] (fp16, gelu fp16 or fp32)."""
    epi = {"f16": 0, "gelu": 1, "f32": 4}[epilogue]
    a, w = a.contiguous(), w.contiguous()
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.is_cuda
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32 if epi == 4 else torch.float16, device=a.device)
    b = bias.float().contiguous() if bias is not None else None
    with torch.cuda.device(a.device):
        check(lib().b200_linear(C.c_void_p(a.data_ptr()), K, C.c_void_p(w.data_ptr()), K,
                                C.c_void_p(b.data_ptr()) if b is not None else None, M, N, K, epi,
                                C.c_void_p(out.data_ptr()), N, block_n, _stream_ptr()))
    return out


def kernel_launches():
    return int(lib().b200_kernel_launches())


PROFILE_CATEGORIES = ("gemm", "attention", "norm", "other", "conv")


def profile_enable(on=True):
    """Per-launch CUDA-event timing by kernel category (disables nothing; graphs simply are not timed)."""
    check(lib().b200_profile_enable(int(bool(on))))


def profile_collect():
    n = len(PROFILE_CATEGORIES)
    ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(lib().b200_profile_collect(ms, fl, by, cnt))
    return {c: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(cnt[i])) for i, c in enumerate(PROFILE_CATEGORIES)}
