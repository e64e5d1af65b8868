"""The oracle restatement against golden vectors produced by the unmodified reference (CPU only)."""
import os

import pytest
import torch

from oracle import m2mixer_oracle as O
from tests.golden_util import BLOCK_KEYS, POS_WEIGHT, block_inputs, full_grad, load, rebuild, rel_err

AV = ["avmnist_S_b8", "avmnist_S_sum_b8", "avmnist_M_b4", "avmnist_B_b16", "avmnist_B_b2"]


def _check(z, out, grads, tol):
    for k in [k for k in z if k.startswith("out.")]:
        assert rel_err(out[k[4:]], z[k]) < tol, k
    for k in [k for k in z if k.startswith("gnorm.")]:
        name = k[6:]
        g = grads[name].double()
        # token_mix.2.net.3.bias has an exactly-zero gradient: a per-token constant over d is removed by
        # every downstream LayerNorm.  Such entries only get an absolute bound.
        if float(z[k]) < 1e-9:
            assert float(g.norm()) < 1e-5, k
            continue
        assert abs(float(g.norm()) - float(z[k])) <= tol * float(z[k]), k
        assert rel_err(g.flatten()[:16], z["ghead." + name]) < max(tol, 1e-5) * 10, k
        if "grad." + name in z:
            assert rel_err(g, z["grad." + name]) < max(tol, 2e-6), k
        if "grad16." + name in z:     # every element, stored as scaled float16 (2^-11 steps)
            assert rel_err(g, full_grad(z, name)) < 5e-4, k


@pytest.mark.parametrize("name", AV)
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_avmnist_matches_reference(name, dtype, tol):
    z, sd, batch = rebuild(name, dtype)
    fusion = "SumFusion" if "sum" in name else "ConcatFusion"
    out = O.avmnist_shared_step(sd, batch, fusion=fusion, training=True)
    _check(z, out, O.grads_of(out["loss"], sd), tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_mimic_matches_reference(dtype, tol):
    z, sd, batch = rebuild("mimic_H_b16", dtype)
    out = O.mimic_shared_step(sd, batch, training=True)
    _check(z, out, O.grads_of(out["loss"], sd), tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_mmimdb_matches_reference(dtype, tol):
    z, sd, batch = rebuild("mmimdb_tiny_b6", dtype)
    out = O.mmimdb_shared_step(sd, batch, torch.tensor(POS_WEIGHT, dtype=dtype), text_encoder="PNLPMixer")
    _check(z, out, O.grads_of(out["loss"], sd), tol)


@pytest.mark.parametrize("name", ["block_odd", "block_b"])
def test_mixer_block_matches_reference(name):
    from oracle.seeding import seeded_state_dict
    z = load(name)
    B, N, D, T, C = (int(v) for v in z["meta.dims"])
    shapes = {"token_mix.0.weight": (D,), "token_mix.0.bias": (D,), "token_mix.2.net.0.weight": (T, N),
              "token_mix.2.net.0.bias": (T,), "token_mix.2.net.3.weight": (N, T), "token_mix.2.net.3.bias": (N,),
              "channel_mix.0.weight": (D,), "channel_mix.0.bias": (D,), "channel_mix.1.net.0.weight": (C, D),
              "channel_mix.1.net.0.bias": (C,), "channel_mix.1.net.3.weight": (D, C), "channel_mix.1.net.3.bias": (D,)}
    sd = {k: v.requires_grad_(True) for k, v in seeded_state_dict(shapes, 77, torch.float64).items()}
    x = torch.tensor(z["x"], requires_grad=True)
    y = O.mixer_block(x, sd, "")
    assert rel_err(y, z["y"]) < 1e-12
    gs = torch.autograd.grad(y, [x] + list(sd.values()), torch.tensor(z["dy"]))
    assert rel_err(gs[0], z["dx"]) < 1e-11
    for k, g in zip(sd, gs[1:]):
        assert abs(float(g.norm()) - float(z["gnorm." + k])) < 1e-10 * (1 + float(z["gnorm." + k]))
        if "grad." + k in z:
            assert rel_err(g, z["grad." + k]) < 1e-11


@pytest.mark.parametrize("name", ["block_c5", "block_c4text", "block_c4fus"])
def test_large_config_blocks_match_reference(name):
    """One MixerBlock at the C4 / C5 layer shapes (BASELINE configs 4, 5; reference modules/mixer.py:25-47, 232-264)."""
    from oracle.seeding import seeded_state_dict
    z = load(name)
    B, N, D, T, C = (int(v) for v in z["meta.dims"])
    sd = {k: v.requires_grad_(True) for k, v in seeded_state_dict(BLOCK_KEYS(N, D, T, C), 77, torch.float64).items()}
    x, dy = block_inputs(z)
    x.requires_grad_(True)
    y = O.mixer_block(x, sd, "")
    assert abs(float(y.norm()) - float(z["ynorm"])) < 1e-10 * float(z["ynorm"])
    assert rel_err(y, z["y"]) < 2e-7          # stored as float32
    gs = torch.autograd.grad(y, [x] + list(sd.values()), dy)
    assert abs(float(gs[0].norm()) - float(z["dxnorm"])) < 1e-10 * float(z["dxnorm"])
    assert rel_err(gs[0], z["dx"]) < 2e-7
    for k, g in zip(sd, gs[1:]):
        gn = float(z["gnorm." + k])
        assert abs(float(g.norm()) - gn) < 1e-10 * (1 + gn), k
        samp = g.flatten()[::int(z["gstride." + k])][:4096]
        assert float((samp - torch.as_tensor(z["gsamp." + k])).abs().max()) < 1e-10 * (1 + gn), k


def test_concat_shape_contract():
    # tests/modules/test_fusion.py:14-24 and :38-47 of the reference (shape-only pins)
    a = torch.ones(10, 20, 30)
    assert O.concat_fusion(a, a).shape == (10, 40, 30)
    assert O.concat_out_shape(20, 20, dim=1) == 40
    assert O.concat_out_shape(a.shape, a.shape) == (10, 40, 30)
    with pytest.raises(ValueError):
        O.concat_out_shape(a.shape, a.shape, dim=2)
    assert O.sum_fusion(a, a).shape == (10, 20, 30)


@pytest.mark.skipif(not os.path.isdir("/root/reference/modules"), reason="reference mount absent (GPU box)")
def test_oracle_against_live_reference():
    import sys
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        import modules as ref
    finally:
        sys.path.remove("/root/reference")
    torch.manual_seed(3)
    m = ref.MLPMixer(in_channels=2, hidden_dim=48, patch_size=4, image_size=[8, 12], num_mixers=2, token_dim=9,
                     channel_dim=50).double()
    x = torch.randn(3, 2, 8, 12, dtype=torch.float64)
    sd = {"m." + k: v for k, v in m.state_dict().items()}
    assert rel_err(O.mlp_mixer(x, sd, "m."), m(x)) < 1e-12
    m2 = ref.MLPMixerNoPatching(hidden_dim=16, num_patch=5, num_mixers=1, token_dim=4, channel_dim=20, embedding_dim=7,
                                proj_dim=16).double()
    x2 = torch.randn(3, 5, 7, dtype=torch.float64)
    assert rel_err(O.mlp_mixer_no_patching(x2, {"m." + k: v for k, v in m2.state_dict().items()}, "m."), m2(x2)) < 1e-12
    m3 = ref.MLP(input_dim=5, hidden_dim=8, num_blocks=2, output_dim=6).double()
    x3 = torch.randn(4, 5, dtype=torch.float64)
    assert rel_err(O.mlp_encoder(x3, {"m." + k: v for k, v in m3.state_dict().items()}, "m."), m3(x3)) < 1e-12
