"""cfg1 (one 50-phoneme utterance): kernel-level breakdown of one `infer` call under torch.profiler (CUPTI)."""
import json
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import bench  # noqa: E402
from baseline import reference_arm  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
model, _ = bench.build_models(dev)
ph, _, cls = reference_arm._inputs(11, 1, 50, 51)
ph, cls = ph.to(dev), cls.to(dev)


def step():
    torch.manual_seed(7)
    return model.infer(ph, style_prompt=cls, use_max=True, noise_scale=0.5)


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
with tempfile.TemporaryDirectory() as td:
    path = Path(td) / "trace.json"
    prof.export_chrome_trace(str(path))
    ev = json.loads(path.read_text())["traceEvents"]
ks = sorted(((e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev
             if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e), key=lambda t: t[0])
span = ks[-1][1] - ks[0][0]
busy = sum(b - a for a, b, _ in ks)
print(f"cfg1: device span {span / 1e3:.2f} ms, {len(ks)} kernels, busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms")
agg = defaultdict(lambda: [0, 0.0])
for a, b, n in ks:
    n = n.split("(")[0][-60:]
    agg[n][0] += 1
    agg[n][1] += b - a
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {t / 1e3:8.2f} ms  {c:5d} x {t / c:8.1f} us  {n}")
