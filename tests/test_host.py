"""CPU-side tests: C-ABI surface, drop-in module contracts, config plumbing.  No GPU, no compute calls."""
import ctypes
import os
import re
import sys

import pytest
import torch

from tests.golden_util import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "m2b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m2b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from m2_mixer_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/m2b200.h but not exported by libm2b200.so"
    assert set(_lib.PROTOTYPES) == set(names), "ctypes prototypes and the header disagree"
    assert lib.m2b200_abi_version() == 1
    assert lib.m2b200_status_string(0) == b"ok"
    # workspace queries are pure host arithmetic
    assert lib.m2b200_channel_mix_workspace_bytes(16384, 128, 3072, 1, 0) == 0          # fused forward: nothing
    # fused backward (generation 4): bf16 LN(u) and dY operand copies + the spilled dH, chunk-major [ceil(C / 64)][M][64]
    assert lib.m2b200_channel_mix_workspace_bytes(16384, 128, 3072, 1, 1) == 2 * 16384 * 128 * 2 + 16384 * 3072 * 2
    assert lib.m2b200_channel_mix_workspace_bytes(256, 128, 3078, 1, 1) == 2 * 256 * 128 * 2 + 256 * 49 * 64 * 2
    assert lib.m2b200_channel_mix_workspace_bytes(12544, 768, 3072, 1, 1) > 2 * 12544 * 3072 * 2   # unfused D > 128: G / dH spill
    assert lib.m2b200_channel_mix_workspace_bytes(100, 128, 64, 0, 0) >= 100 * (128 + 64) * 4


def test_ctypes_prototypes_match_header_signatures():
    """Argument count and kind (pointer / int / int64 / float / size_t / uint64) of every ctypes prototype against the
    parameter list parsed from include/m2b200.h - a mismatch here is a silent ABI break on the GPU box."""
    import ctypes as C
    from m2_mixer_b200 import _lib
    src = open(os.path.join(ROOT, "include", "m2b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    kinds = {C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_int: "int", C.c_int64: "i64", C.c_float: "f32", C.c_size_t: "size",
             C.c_uint64: "u64", C.c_ulonglong: "u64"}

    def kind_of_decl(d):
        d = d.strip()
        if "*" in d:
            return "ptr"
        t = d.rsplit(" ", 1)[0].replace("const", "").strip()
        return {"int": "int", "int64_t": "i64", "float": "f32", "size_t": "size", "uint64_t": "u64"}[t]

    for name, (_res, args) in _lib.PROTOTYPES.items():
        m = re.search(r"\b%s\s*\(([^)]*)\)" % name, src)
        assert m, name
        decl = m.group(1).strip()
        params = [] if decl in ("", "void") else [x for x in decl.split(",")]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes prototype {len(args)}"
        for i, (d, a) in enumerate(zip(params, args)):
            k = kinds.get(a, "ptr")     # POINTER(...) types are pointers
            same64 = {"size", "u64"}     # c_size_t and c_uint64 are one ctypes type on LP64
            assert kind_of_decl(d) == k or {kind_of_decl(d), k} <= same64, f"{name} arg {i}: header `{d.strip()}` vs ctypes {a}"


def test_ops_fail_loudly_without_cuda():
    from m2_mixer_b200 import functional as F
    x = torch.zeros(2, 4, 8)
    w = torch.ones(8)
    with pytest.raises((NotImplementedError, RuntimeError)):
        F.layer_norm(x, w, w)


@pytest.mark.parametrize("golden,preset", [("avmnist_S_b8", "avmnist_S"), ("avmnist_M_b4", "avmnist_M"),
                                           ("avmnist_B_b16", "avmnist_B"), ("mimic_H_b16", "mimic_H"),
                                           ("mmimdb_tiny_b6", "mmimdb_tiny")])
def test_state_dict_layout_matches_reference(golden, preset):
    """Key names AND shapes recorded from the reference's own modules (tests/golden/make_golden.py)."""
    from m2_mixer_b200 import models, presets
    cfg = presets.get(preset)
    m = models.get_model(cfg["type"])(cfg, {})
    z = load(golden)
    ref = {k: tuple(int(t) for t in s.split(",")) for k, s in zip(z["meta.keys"].tolist(), z["meta.shapes"].tolist())}
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ours == ref
    assert all(v.dtype == torch.float32 for v in m.state_dict().values())


def test_reference_ckpt_style_roundtrip(tmp_path):
    from m2_mixer_b200 import models, presets
    from oracle.seeding import seeded_state_dict
    cfg = presets.get("avmnist_S")
    m = models.AVMnistMixerMultiLoss(cfg, dict(presets.AVMNIST_OPTIM))
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5)
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": sd, "epoch": 3}, path)          # Lightning checkpoints carry no hparams
    m2 = models.AVMnistMixerMultiLoss.load_from_checkpoint(str(path), model_cfg=cfg, optimizer_cfg=dict(presets.AVMNIST_OPTIM))
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k])
    assert m2.scheduler_patience == 2 and "scheduler_patience" not in m2.optimizer_cfg


def test_constructor_contracts():
    from m2_mixer_b200 import modules as M
    blk = M.MLPMixer(in_channels=1, hidden_dim=32, patch_size=14, image_size=[28, 28], num_mixers=2, token_dim=16,
                     channel_dim=256, dropout=0.1, block_type="MLPMixer", some_unknown_key=1)      # **kwargs swallowed
    assert blk.num_patch == 4 and len(blk.mixer_blocks) == 2
    with pytest.raises(AssertionError):
        M.MLPMixer(in_channels=1, hidden_dim=32, patch_size=5, image_size=[28, 28], num_mixers=1, token_dim=4, channel_dim=8)
    assert M.get_block_by_name(block_type="FusionMixer", hidden_dim=32, num_patches=8, num_mixers=1, token_dim=4,
                               channel_dim=8, fusion_function="ConcatFusion").num_patch == 8
    assert isinstance(M.get_fusion_by_name(fusion_function="SumFusion", useless_arg=3), M.SumFusion)
    assert M.get_classifier_by_name(classifier="StandardClassifier", input_shape=[16, 49, 32],
                                    num_classes=10).classifer.weight.shape == (10, 32)


def test_fusion_shape_contracts():
    """The pins of the reference's own tests (tests/modules/test_fusion.py:14-24, 38-47)."""
    from m2_mixer_b200.modules import ConcatFusion, SumFusion
    f = ConcatFusion(dim=1, useless_arg=None)
    assert f.get_output_shape(20, 20, dim=1) == 40
    assert f.get_output_shape(20, 20, dim=2) == 20
    assert f.get_output_shape((10, 20, 30), (10, 20, 30)) == (10, 40, 30)
    with pytest.raises(ValueError):
        f.get_output_shape(torch.ones(10, 20, 30), torch.ones(10, 20, 30), dim=2)
    s = SumFusion(useless_arg=None)
    assert s.get_output_shape(20, 20, dim=1) == 20
    assert s.get_output_shape((10, 20, 30), (10, 20, 30)) == (10, 20, 30)
    with pytest.raises(ValueError):
        s.get_output_shape(20, 30, dim=1)
    with pytest.raises(ValueError):
        s.get_output_shape(torch.ones(10, 20, 30), torch.ones(10, 20, 30), dim=2)


def test_yaml_loader_coerces_and_overrides(tmp_path):
    from m2_mixer_b200.config import deep_update, load_yaml
    p = tmp_path / "c.yml"
    p.write_text("train:\n  optimizer:\n    lr: 1e-2\n    betas: [0.9, 0.999]\n    eps: 1e-8\nmodel:\n  type: X\n  dropout: 0.5\n")
    c = load_yaml(str(p))
    assert isinstance(c.train.optimizer.lr, float) and c.train.optimizer.lr == 1e-2 and c.train.optimizer.eps == 1e-8
    assert c.model.get("missing", 7) == 7 and c.model.type == "X"
    deep_update(c, "model.modalities.image.hidden_dim", 64)
    assert c.model.modalities.image.hidden_dim == 64


def test_dropout_argument_validation():
    from m2_mixer_b200 import modules as M
    with pytest.raises(ValueError):
        M.MixerBlock(32, 4, 8, 64, dropout=1.0)
    blk = M.MixerBlock(32, 4, 8, 64, dropout=0.5)
    assert blk.dropout_p == 0.5 and blk.token_mix[2].dropout_p == 0.5


def test_max_and_mean_fusion_shape_contracts_match_reference():
    """MaxFusion / MeanFusion (reference modules/fusion.py:190-204, 258-272): same get_output_shape answers and errors."""
    from m2_mixer_b200 import modules as M
    for cls in (M.MaxFusion, M.MeanFusion):
        f = cls()
        assert f.get_output_shape((4, 8, 32), (4, 8, 32)) == (4, 8, 32)
        assert f.get_output_shape(8, 8, dim=1) == 8
        with pytest.raises(ValueError):
            f.get_output_shape((4, 8, 32), (4, 9, 32))
        with pytest.raises(ValueError):
            f.get_output_shape((4, 8, 32), (4, 8, 32), dim=1)
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):
        sys.path.insert(0, ref_root)
        try:
            import importlib
            rf = importlib.import_module("modules.fusion")
            for name in ("MaxFusion", "MeanFusion"):
                a, b = getattr(rf, name)(), getattr(M, name)()
                assert a.get_output_shape((2, 3, 4), (2, 3, 4)) == b.get_output_shape((2, 3, 4), (2, 3, 4))
                assert a.get_output_shape(5, 5, dim=1) == b.get_output_shape(5, 5, dim=1)
        finally:
            sys.path.remove(ref_root)
            for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
                del sys.modules[k]


def test_bench_reference_arm_emits_one_json_line():
    """`bench.py --impl reference` (the reference's algorithm on the host cores: the oracle port, no GPU involved) prints exactly
    one JSON line with the contract's keys."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-batch", "8"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("train samples/s") and d["unit"] == "samples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/cfg"), reason="reference mount absent (GPU box)")
@pytest.mark.parametrize("rel", ["cfg/avmnist/avmnist_m2-mixer_S.yml", "cfg/avmnist/avmnist_m2-mixer_M.yml",
                                 "cfg/avmnist/avmnist_m2-mixer_B.yml", "cfg/mimic/mimic_m2-mixer_H.yml"])
def test_reference_yaml_files_build_the_drop_in_models(rel):
    """The reference's own cfg/*.yml (run.py:28: OmegaConf.load) through config.load_yaml into the drop-in task modules:
    same class name as cfg model.type, and the SAME state-dict keys / shapes as the reference modules assembled from the
    same file (reference models/avmnist.py:178-196, models/mimic.py:36-52 via tests/golden/make_golden.py)."""
    import importlib
    from m2_mixer_b200 import models
    from m2_mixer_b200.config import load_yaml
    cfg = load_yaml(os.path.join("/root/reference", rel))
    assert isinstance(cfg.train.optimizer.lr, float)          # '1e-2' style strings are coerced like OmegaConf does
    m = models.get_model(cfg.model.type)(cfg.model, cfg.train.optimizer)
    assert type(m).__name__ == cfg.model.type
    mg = importlib.import_module("tests.golden.make_golden")
    ref = mg.RefMimic(mg.load_cfg(rel)["model"]) if "mimic" in rel else mg.RefAVMnist(mg.load_cfg(rel)["model"])
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert ours == theirs
    ref.load_state_dict(m.state_dict(), strict=True)          # and the tensors move across in both directions
    m.load_state_dict(ref.state_dict(), strict=True)


def test_ops_have_fake_kernels_for_meta_tracing():
    """every m2b200:: op carries a fake (meta) kernel, so FakeTensor / torch.compile tracing sees shapes and
    dtypes without launching anything (runs without a GPU: fake CUDA tensors)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from m2_mixer_b200 import ops  # noqa: F401  registers the ops
    o = torch.ops.m2b200
    with FakeTensorMode():
        f = lambda *s: torch.empty(*s, device="cuda")
        x, ln = f(8, 4, 128), f(128)
        w1, b1, w2, b2 = f(32, 4), f(32), f(4, 32), f(4)
        assert o.token_mix_fwd(x, ln, ln, w1, b1, w2, b2, 1).shape == (8, 4, 128)
        g = o.token_mix_bwd(x, x, ln, ln, w1, b1, w2, 1)
        assert [tuple(t.shape) for t in g] == [(8, 4, 128), (128,), (128,), (32, 4), (32,), (4, 32), (4,)]
        cw1, cb1, cw2 = f(3072, 128), f(3072), f(128, 3072)
        assert o.channel_mix_fwd(x, ln, ln, cw1, cb1, cw2, ln, None, None, 1).shape == (8, 4, 128)
        assert o.channel_mix_bwd(x, x, ln, ln, cw1, cb1, cw2, None, None, 1)[3].shape == (3072, 128)
        assert o.patch_embed_fwd(f(8, 1, 112, 112), f(128, 1, 56, 56), None, None, 56, 1).shape == (8, 4, 128)
        assert o.layernorm_concat_fwd([x, x], [ln, ln], [ln, ln]).shape == (8, 8, 128)
        losses, logits, preds = o.heads_loss_fwd([x, x, x], [f(10, 128)] * 3, [f(10)] * 3, torch.empty(8, dtype=torch.int64, device="cuda"),
                                                 None, [1.0, 1.0, 1.0], 0)
        assert losses.shape == (4,) and logits.shape == (3, 8, 10) and preds.dtype == torch.int64


def test_channel_mix_workspace_query_follows_the_backward_generation():
    """The workspace query is the contract between the torch shim and the C ABI: with the default backward (dH spilled through
    TMA stores) it includes the chunk-major dH buffer, with M2B200_CHAIN_GEN=2 (weight gradients recompute G and dH on chip)
    only the two bf16 operand copies.  The generation is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from m2_mixer_b200 import _lib; lib = _lib.load(); "
            "print(lib.m2b200_channel_mix_workspace_bytes(16384, 128, 3072, 1, 1))")
    out = {}
    for gen in ("2", "4"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root,
                           env=dict(os.environ, M2B200_CHAIN_GEN=gen), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[gen] = int(r.stdout.strip().splitlines()[-1])
    assert out["2"] == 2 * 16384 * 128 * 2
    assert out["4"] == out["2"] + 16384 * 3072 * 2
