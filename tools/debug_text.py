"""GPU: where does the cfg2 text side differ from the reference golden?"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from golden_cases import TEXT_CASES, acoustic_inputs, golden_noise
from promptttspp_b200.models.prompttts_mdn_v2_final.model import InferNoise
from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
from promptttspp_b200.utils.synthetic import build_acoustic, synthetic_state_dict
torch.set_grad_enabled(False)
case = TEXT_CASES["cfg2_text"]
gold = {k: torch.from_numpy(v) for k, v in np.load(ROOT / "tests/golden/text_cfg2_text.npz").items()}
phoneme, lengths, cls_emb = acoustic_inputs(case)
model = build_acoustic(rel_pos_type=case["rel_pos_type"], bert=FixedPromptEmbedding(cls_emb), K_step=1)
model.load_state_dict(synthetic_state_dict(model, seed=case["weight_seed"], frames_per_phoneme=case["frames_per_phoneme"]), strict=True)
model = model.cuda().eval()
B = phoneme.shape[0]
z_style = golden_noise(case, B, None).z_style
model.infer_batch(phoneme.cuda(), lengths.cuda(), style_prompt=["p"] * B, use_max=True, noise_scale=case["noise_scale"], noise=InferNoise(z_style, None, None))
d = model.last_durations.cpu(); ld = model.last_log_durations.cpu()
gd = gold["duration"]; gl = gold["log_d"].squeeze(1)
valid = torch.arange(phoneme.shape[1])[None] < lengths[:, None]
err = (ld - gl).abs()
print("log_d max err valid", float(err[valid].max()), "padded", float(err[~valid].max()) if (~valid).any() else 0)
for b, i in (d != gd).nonzero().tolist():
    print("diff at", b, i, "len", int(lengths[b]), "ours", int(d[b, i]), "gold", int(gd[b, i]), "log_d ours", float(ld[b, i]), "gold", float(gl[b, i]),
          "exp", float(ld[b, i].double().exp()), float(gl[b, i].double().exp()))
top = err.masked_fill(~valid, 0).flatten().topk(5)
for v, idx in zip(top.values.tolist(), top.indices.tolist()):
    b, i = divmod(idx, phoneme.shape[1]); print("top err", v, "at", b, i, "len", int(lengths[b]))
