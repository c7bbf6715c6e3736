"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/pttspp_b200.h declares;
the host shims keep the reference's checkpoint contract and refuse to run without CUDA (no fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def libpath():
    from promptttspp_b200 import build

    return build.build()


def declared_symbols():
    text = (ROOT / "include" / "pttspp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pttspp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(libpath):
    from promptttspp_b200 import _abi

    handle = ctypes.CDLL(str(libpath))
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in _abi.SIGNATURES, f"{name} has no ctypes signature in _abi.py"
    assert set(_abi.SIGNATURES) <= set(names)
    assert _abi.lib().pttspp_abi_version() == 1
    assert _abi.lib().pttspp_last_error() is not None


def test_struct_sizes_match_header(libpath):
    """Compile a tiny C probe against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess
    import tempfile

    from promptttspp_b200 import _abi

    src = '#include <stdio.h>\n#include "pttspp_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(pttspp_conv1d_desc),sizeof(pttspp_layernorm_desc),sizeof(pttspp_bigvgan_config)," \
          "sizeof(pttspp_acoustic_config),sizeof(pttspp_diffnet_layer),sizeof(pttspp_diffnet_run_desc));return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "p.c").write_text(src)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(Path(d) / "p.c"), "-o", str(Path(d) / "p")])
        sizes = [int(x) for x in subprocess.check_output([str(Path(d) / "p")]).split()]
    mirrors = [_abi.Conv1dDesc, _abi.LayerNormDesc, _abi.BigVGANConfig, _abi.AcousticConfig, _abi.DiffNetLayer,
               _abi.DiffNetRunDesc]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_host_weight_packing_matches_torch(libpath):
    from promptttspp_b200 import ops

    g = torch.Generator().manual_seed(0)
    v = torch.randn(6, 16, 3, generator=g)
    gg = torch.rand(6, 1, 1, generator=g) + 0.5
    w = torch._weight_norm(v, gg, 0)
    packed = ops.pack_conv_weight(v, gg)
    assert packed.shape == (3, 16, 8)
    assert torch.allclose(packed[:, :, :6], w.permute(2, 1, 0), atol=1e-6)
    assert torch.count_nonzero(packed[:, :, 6:]) == 0
    inter = ops.pack_conv_weight(v, None, interleave_halves=True)
    assert torch.equal(inter[:, :, 0:6:2], v[:3].permute(2, 1, 0)) and torch.equal(inter[:, :, 1:6:2], v[3:].permute(2, 1, 0))
    # transposed conv: phase r, slot k' holds kernel tap r + (J-1-k')*stride
    vt = torch.randn(16, 5, 6, generator=g)
    gt = torch.rand(16, 1, 1, generator=g) + 0.5
    wt = torch._weight_norm(vt, gt, 0)
    pt = ops.pack_convtr_weight(vt, 3, gt)
    assert pt.shape == (3, 2, 16, 8)
    for r in range(3):
        for kp in range(2):
            assert torch.allclose(pt[r, kp, :, :5], wt[:, :, r + (1 - kp) * 3], atol=1e-6)


def test_state_dict_contract_and_no_cpu_fallback(libpath):
    from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
    from promptttspp_b200.utils.model import remove_weight_norm_
    from promptttspp_b200.utils.synthetic import build_acoustic, build_vocoder, synthetic_state_dict

    voc = build_vocoder()
    sd = synthetic_state_dict(voc, seed=1)
    assert len(sd) == 453  # vocoder checkpoint keys
    voc.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        voc(torch.zeros(1, 80, 4))
    voc.apply(remove_weight_norm_)  # synthesize.py:116 does this; must be tolerated
    assert "conv_pre.weight" in voc.state_dict() and "conv_pre.weight_g" not in voc.state_dict()

    ac = build_acoustic(bert=FixedPromptEmbedding(torch.zeros(1, 768)))
    assert len(ac.state_dict()) == 470  # non-BERT acoustic keys
    ac.load_state_dict(synthetic_state_dict(ac, seed=2), strict=True)
    ac.apply(remove_weight_norm_)
    with pytest.raises(RuntimeError, match="CUDA"):
        ac.infer(torch.ones(1, 5, dtype=torch.long), style_prompt=["x"])
    with pytest.raises(AssertionError):
        ac.infer(torch.ones(1, 5, dtype=torch.long))  # neither style input (model.py:209)
    with pytest.raises(NotImplementedError):
        ac.forward(None)
