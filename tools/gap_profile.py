"""Where does the wall time of one cfg2 acoustic step go?  Runs the step under torch.profiler (CUPTI kernel records) and
reports: span of the step on the device, sum of kernel durations, idle time between kernels and the largest gaps with
the kernels around them.  (Evidence for the `wall - sum(kernel)` item of VERDICT r1.)"""
import json
import sys
import tempfile
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import bench  # noqa: E402

torch.set_grad_enabled(False)


def main():
    dev = torch.device("cuda", 0)
    model, _ = bench.build_models(dev)
    ph, ln, cls = bench.cfg2_inputs(seed=2)
    ph, ln, cls = ph.to(dev), ln.to(dev), cls.to(dev)

    def step():
        torch.manual_seed(1000)
        return model.infer_batch(ph, ln, style_prompt=cls, use_max=True, noise_scale=0.5, return_f0=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

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
    gaps = sorted(((ks[i + 1][0] - ks[i][1], ks[i][2][:60], ks[i + 1][2][:60]) for i in range(len(ks) - 1)), reverse=True)
    idle = sum(g for g, _, _ in gaps if g > 0)
    print(f"device span of one step: {span / 1e3:.2f} ms; {len(ks)} kernels/copies, sum of durations {busy / 1e3:.2f} ms, "
          f"idle between them {idle / 1e3:.2f} ms ({100 * idle / span:.2f} % of the span)")
    print("largest gaps (us): ")
    for g, a, b in gaps[:8]:
        print(f"  {g:9.1f}  after {a}  before {b}")
    small = [g for g, _, _ in gaps if 0 < g < 50]
    print(f"gaps below 50 us: {len(small)}, {sum(small) / 1e3:.2f} ms in total, mean {sum(small) / max(1, len(small)):.2f} us")


if __name__ == "__main__":
    main()
