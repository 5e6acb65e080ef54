"""Host side of the APT discriminator forward, mirroring
`WanAPTDiscriminator.forward` (seaweed_apt/model.py:123-186):

    logit = disc(x, t, context, seq_len)                      # [B, 1]
    logit, feats = disc(x, t, context, seq_len, return_features=True)

The reference runs its deep-copied backbone under `torch.no_grad()` with forward hooks on blocks[15], [25], [35]
(model.py:150-163): here that is one `DitEngine.forward` with three residual-stream taps.  The three single-query
heads, the concat and `final_proj` run in libb200dit.so (`b200disc_forward`, csrc/disc_engine.cu).
Inference only, like the rest of the engine: the reference trains the heads through autograd.
"""
import ctypes as C

import torch

from ._lib import B200Error, MAX_ITEMS, check, lib, ptr_array
from .engine import DitEngine, _load_state, _stream_ptr

REFERENCE_TAPS = (16, 26, 36)                    # 1-based block numbers hooked by the reference, model.py:150-155


def timestep_shift(t, frames):
    """model.py:158-159: t' = s t / (1 + (s - 1) t), s = 1 for single-frame latents, 12 for videos."""
    s = 1.0 if frames == 1 else 12.0
    return s * t / (1.0 + (s - 1.0) * t)


class AptDiscriminator:
    """`backbone`: the DitEngine holding the discriminator's copy of the Wan weights (model.py:91).
    `state_dict`: WanAPTDiscriminator.state_dict(); `backbone.*` entries are ignored here.
    `tap_blocks`: 1-based numbers of the blocks whose outputs feed cross_attn_16 / _26 / _36, in that order.  The
    reference hard-codes (16, 26, 36) and raises IndexError on a backbone with fewer than 36 blocks (the 30-block
    1.3B); so does the default here -- pass e.g. (10, 20, 30) explicitly for such a backbone."""

    def __init__(self, backbone, state_dict, tap_blocks=REFERENCE_TAPS, qk_norm=True):
        if not isinstance(backbone, DitEngine):
            raise TypeError("backbone must be a b200dit.DitEngine")
        if len(tap_blocks) != 3:
            raise ValueError("the discriminator has exactly three heads")
        for b in tap_blocks:
            if not 1 <= b <= backbone.cfg["num_layers"]:
                raise IndexError(f"index {b - 1} is out of range")        # nn.ModuleList's message, model.py:154
        self.backbone, self.tap_blocks = backbone, tuple(int(b) for b in tap_blocks)
        self.device, self.dim = backbone.device, backbone.cfg["dim"]
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().b200disc_create(self.dim, backbone.cfg["num_heads"], int(bool(qk_norm)),
                                        float(backbone.cfg["eps"]), C.byref(self._h)))
            keep = _load_state(lib().b200disc_load_weight, self._h, state_dict, lambda n: not n.startswith("backbone."))
            check(lib().b200disc_finalize(self._h))
        self._taps, self._rows = None, 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().b200disc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def heads(self, taps, n_items, rows_per_item, return_features=False):
        """model.py:166-186 on given block outputs: three fp32 [n_items * rows_per_item, dim] device tensors."""
        for u in taps:
            if u.dtype != torch.float32 or not u.is_cuda or not u.is_contiguous() or u.numel() != n_items * rows_per_item * self.dim:
                raise B200Error("taps must be contiguous fp32 device tensors of n_items * rows_per_item * dim elements")
        logits = torch.empty((n_items, 1), dtype=torch.float32, device=self.device)
        feats = torch.empty((3, n_items, 1, self.dim), dtype=torch.float32, device=self.device) if return_features else None
        with torch.cuda.device(self.device):
            check(lib().b200disc_forward(self._h, ptr_array([u.data_ptr() for u in taps]), n_items, rows_per_item,
                                         C.c_void_p(logits.data_ptr()),
                                         C.c_void_p(feats.data_ptr() if feats is not None else 0), _stream_ptr()))
        return (logits, [feats[0], feats[1], feats[2]]) if return_features else logits

    @torch.no_grad()
    def forward(self, x, t, context, seq_len, return_features=False):
        """x: [B, C, T, H, W] (or a list of [C, T, H, W] of one shape); t: [B]; context: list of [rows, text_dim]."""
        xs = [u for u in x]
        n = len(xs)
        _, T, H, W = xs[0].shape
        if any(u.shape != xs[0].shape for u in xs):
            raise B200Error("all items of a discriminator batch must share one latent shape")
        L = T * (H // 2) * (W // 2)
        if seq_len < L:
            raise AssertionError(f"Max seq len {L} exceeds limit {seq_len}")                 # wan model.py:521
        # seq_len > L: the reference's block outputs are [B, seq_len, dim] and its heads attend over the padded rows
        # too (seaweed_apt/model.py:162-171; wan model.py:522): the backbone carries them for this call
        R = int(seq_len)
        ts = timestep_shift(torch.as_tensor(t, dtype=torch.float32).reshape(-1), T)
        # one backbone call fills the taps of the items it co-batches (engine.py: at most MAX_ITEMS, one item
        # when the per-item row count is not a multiple of 8): run the heads once per such call
        step = MAX_ITEMS if (L % 8 == 0 and R % 8 == 0) else 1
        logits, feats = [], []
        for s0 in range(0, n, step):
            part = range(s0, min(n, s0 + step))
            m = len(part)
            reuse = self._taps if self._rows == m * R else None
            self._taps = self.backbone.set_taps([b - 1 for b in self.tap_blocks], m * R, buffers=reuse)
            self._rows = m * R
            self.backbone.set_pad_to_seq_len(R > L)
            try:
                self.backbone.forward([xs[i] for i in part], ts[s0:s0 + m], [context[i] for i in part], seq_len)
            finally:
                self.backbone.set_tap(None)
                self.backbone.set_pad_to_seq_len(False)
            r = self.heads(self._taps, m, R, return_features)
            logits.append(r[0] if return_features else r)
            if return_features:
                feats.append(r[1])
        logit = logits[0] if len(logits) == 1 else torch.cat(logits)
        if not return_features:
            return logit
        return logit, [torch.cat([g[k] for g in feats]) for k in range(3)]

    __call__ = forward
