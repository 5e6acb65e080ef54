"""The drop-in boundary against the REAL reference classes, on CPU (container-only: needs /root/reference).

`DitEngine.from_module` / `install` read a live `WanModel`'s attributes and state_dict, `install_vae` a live
`WanVAE_`'s state_dict.  Round 1 exercised them on stand-in classes only; here the unmodified classes are built
through oracle/ref_loader.py and the host-side extraction runs for real: the architecture read off the module
equals the constructor arguments, and the module's state_dict keys / element counts equal what an engine of that
architecture loads (b200dit_weight_names is host-only, so no GPU is needed)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference tree not mounted")

TINY = dict(dim=128, ffn_dim=256, num_heads=1, num_layers=2, text_dim=32, in_dim=16, out_dim=16, freq_dim=256)


@pytest.mark.parametrize("i2v", [False, True])
def test_from_module_extraction_on_real_wanmodel(i2v):
    import b200dit
    M, _ = ref_loader.load_reference_modules()
    kw = dict(TINY, in_dim=32 if i2v else 16)
    m = M.WanModel(model_type="i2v" if i2v else "t2v", use_checkpoint=False, **kw).eval()
    cfg = b200dit.DitEngine.config_from_module(m)
    for k, v in kw.items():
        assert cfg[k] == v, k
    assert cfg["i2v"] == i2v and cfg["text_len"] == 512 and cfg["eps"] == 1e-6
    want = b200dit.DitEngine.expected_weight_names(**cfg)
    have = {k: v.numel() for k, v in m.state_dict().items() if k != "freqs"}     # engine.load_state_dict's filter
    assert set(have) == set(want), (set(have) ^ set(want))
    assert have == want


def test_real_1p3b_architecture_names():
    """The 1.3B configuration (configs/wan_t2v_1_3B.py:20-29) on the meta device: 30 blocks, 1.419 B parameters."""
    import b200dit
    M, _ = ref_loader.load_reference_modules()
    with torch.device("meta"):
        m = M.WanModel(model_type="t2v", dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, use_checkpoint=False)
    cfg = b200dit.DitEngine.config_from_module(m)
    want = b200dit.DitEngine.expected_weight_names(**cfg)
    have = {k: v.numel() for k, v in m.state_dict().items() if k != "freqs"}
    assert have == want
    assert abs(sum(want.values()) - 1.419e9) < 2e6


def test_install_dispatch_under_autograd_on_real_wanmodel():
    """wan_shim.install's dispatch on the real class without a GPU: under autograd a plain t2v call (the student step,
    distilled_trainer.py:268-301) goes to the engine's autograd node; what the engine's backward does not cover
    (contexts that require gradients, `y` / `clip_fea`) falls through to the ORIGINAL forward; the weight signature
    moves with an in-place update."""
    import b200dit
    from b200dit import wan_shim
    M, _ = ref_loader.load_reference_modules()
    m = M.WanModel(model_type="t2v", use_checkpoint=False, **TINY).train()

    class Taken(Exception):
        pass

    class NoEngine:                                  # no GPU here: records which path the shim takes
        cfg = dict(i2v=0, out_dim=16)
        device = torch.device("cpu")

        def forward(self, *a, **k):
            raise AssertionError("inference path taken under autograd")

        def _t_tensor(self, t, n):
            return torch.as_tensor(t, dtype=torch.float32).reshape(-1)

        def train_forward(self, *a, **k):
            raise Taken()

        def load_state_dict(self, sd):
            self.reloaded = True

    eng = NoEngine()
    b200dit.install(m, engine=eng)
    x = [torch.randn(16, 1, 4, 4)]
    with pytest.raises(Taken):
        m(x, t=torch.tensor([500.0]), context=[torch.randn(5, 32)], seq_len=4)
    out = m(x, t=torch.tensor([500.0]), context=[torch.randn(5, 32, requires_grad=True)], seq_len=4)
    assert out[0].shape == (16, 1, 4, 4) and out[0].requires_grad        # the reference's own graph
    s0 = wan_shim._weights_signature(m)
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    assert wan_shim._weights_signature(m) != s0
    b200dit.uninstall(m)


def test_install_vae_extraction_on_real_wanvae():
    import b200dit
    _, V = ref_loader.load_reference_modules()
    vae = V.WanVAE_(dim=8, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                    temperal_downsample=[False, True, True], dropout=0.0)
    sd = vae.state_dict()
    assert b200dit.VaeEngine.config_from_state_dict(sd) == dict(dim=8, z_dim=16)
    # the prefixes VaeEngine.load_state_dict forwards cover the whole module
    assert all(k.startswith(("decoder.", "conv2.", "encoder.", "conv1.")) for k in sd)
