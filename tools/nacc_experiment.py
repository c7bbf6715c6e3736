"""cfg2-scale A/B of the one-accumulator mode for ALL CTA-pair launches (PTTSPP_UMMA_NACC=1): mel difference and time."""
import os, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
torch.set_grad_enabled(False)
dev = torch.device("cuda")
model, _ = bench.build_models(dev)
ph, ln, cls = [t.to(dev) for t in bench.cfg2_inputs(seed=2)]
def run():
    torch.manual_seed(1000)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mel, cf0, vuv, fl = model.infer_batch(ph, ln, style_prompt=cls, use_max=True, noise_scale=0.5, return_f0=True)
    torch.cuda.synchronize()
    return mel, (time.perf_counter() - t0) * 1e3
run()
m2, t2 = run()
os.environ["PTTSPP_UMMA_NACC"] = "1"
run()
m1, t1 = run()
print(f"two accumulators {t2:.1f} ms, one accumulator {t1:.1f} ms, mel max-abs diff {float((m1 - m2).abs().max()):.3e}, "
      f"rms diff {float((m1 - m2).pow(2).mean().sqrt()):.3e}, mel absmax {float(m2.abs().max()):.2f}")
