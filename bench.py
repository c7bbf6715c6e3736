"""Benchmark of the PromptTTS++ inference hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Headline workload (N=1): BASELINE.json configs[1] -- prompttts_mdn_v2 acoustic model, batch 16 synthetic
phoneme sequences (len <= 256), full text -> mel forward incl. the 100-step diffusion decoder; metric =
valid mel frames per second.  The same JSON line carries the BigVGAN leg of the metric (configs[2]:
16 x 1024-frame mel -> 24 kHz waveform, real-time factor) under "bigvgan".
A "step" is one pass of the path over one batch.  N > 1 (torchrun, one rank per GPU): every rank runs its
own batch (weak scaling, utterances are independent -- no data-path collective); time = max over ranks.
`--impl reference` times the CPU oracle port of the reference (oracle/oracle.py) on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")

import torch  # noqa: E402

WORKLOAD = ("cfg2: prompttts_mdn_v2 acoustic model only, batch 16 synthetic phoneme seqs len<=256 (legacy rel-pos demo "
            "config), text->mel incl. 100-step diffusion")
METRIC = "mel_frames_per_sec"
UNIT = "frames/s"
TAGS = ["conv1d_simt_fp32", "conv1d_tcgen05_splitfp16", "aa_snake", "layernorm", "relpos_attention", "other"]
# algorithmic work per padded mel frame
DIFFNET_FLOP_PER_FRAME = 2.643e9
BIGVGAN_FLOP_PER_FRAME = 422.5e6


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------

def cfg2_inputs(seed, B=16, lo=128, hi=257):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(lo, hi, (B,), generator=g)
    lengths[0] = hi - 1
    Tx = int(lengths.max())
    phoneme = torch.zeros(B, Tx, dtype=torch.int64)
    for b in range(B):
        phoneme[b, : int(lengths[b])] = torch.randint(3, 90, (int(lengths[b]),), generator=g)
    cls_emb = torch.randn(B, 768, generator=g)
    return phoneme, lengths, cls_emb


def cfg3_inputs(seed=3, B=16, T=1024):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


def build_models(device):
    from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
    from promptttspp_b200.utils.synthetic import build_acoustic, build_vocoder, synthetic_state_dict

    model = build_acoustic(bert=FixedPromptEmbedding(torch.zeros(1, 768)))
    model.load_state_dict(synthetic_state_dict(model, seed=1234), strict=True)
    voc = build_vocoder()
    voc.load_state_dict(synthetic_state_dict(voc, seed=4321), strict=True)
    return model.to(device).eval(), voc.to(device).eval()


def prof_report(lib):
    n = 6
    ms, fl, by = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
    calls = (ctypes.c_int64 * n)()
    from promptttspp_b200 import _abi

    _abi.check(lib.pttspp_prof_report(ms, fl, by, calls, n))
    return [dict(kernel=TAGS[i], ms=ms[i], flops=fl[i], bytes=by[i], calls=calls[i]) for i in range(n)]


def ncu_traffic(kernel, rows):
    """DRAM bytes per launch of `kernel`'s family from the committed ncu --set full capture (profiles/), scaled from
    the capture's row count to this run's; None when no capture covers the family."""
    p = ROOT / "profiles" / "r01_ncu_traffic.json"
    if not p.exists():
        return None, None
    t = json.loads(p.read_text())
    if kernel == "conv1d_tcgen05_splitfp16" and "pair_dilated_gate" in t and "pair_dual_1x1" in t and rows:
        per = 0.5 * (t["pair_dilated_gate"]["dram_bytes_per_launch"] + t["pair_dual_1x1"]["dram_bytes_per_launch"])
        return per * rows / t["_rows"], ("profiles/r01_ncu_traffic.json: mean of the two DiffNet CTA-pair kernels "
                                         f"(dilated+gate, dual 1x1), scaled {t['_rows']} -> {int(rows)} rows")
    if kernel == "aa_snake" and "aa_snake" in t:
        return t["aa_snake"]["dram_bytes_per_launch"], "profiles/r01_ncu_traffic.json (captured launch, not rescaled)"
    return None, None


def roofline_from(report, pk, in_long_step, rows=None):
    """Dominant kernel family of one profiled step -> the `roofline` object."""
    top = max(report, key=lambda r: r["ms"])
    if top["ms"] <= 0 or top["calls"] == 0:
        return None
    traffic, traffic_src = ncu_traffic(top["kernel"], rows)
    if top["flops"] > 0:
        peak = pk["tf_sustained"] if in_long_step else pk["tf_burst"]
        ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": top["kernel"], "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "launches": top["calls"],
                "avg_launch_ms": top["ms"] / top["calls"], "share_of_step": top["ms"] / sum(r["ms"] for r in report),
                "mma_frac": (3.0 * ach / peak) if top["kernel"] == TAGS[1] else None,
                "peak_source": pk["source"] + (", bf16 dense sustained" if in_long_step else ", bf16 dense burst"),
                "note": ("fp32 CUDA-core FFMA path; the tensor peak is the contract's denominator"
                         if top["kernel"] == TAGS[0] else
                         "split-fp16 (3 tcgen05 MMAs per fp32 product): ceiling = 1/3 of the bf16 peak; mma_frac = "
                         "issued fp16 MMA FLOPs / peak")}
    ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": top["kernel"], "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
            "frac": ach / pk["hbm"], "traffic": traffic, "traffic_source": traffic_src, "launches": top["calls"],
            "avg_launch_ms": top["ms"] / top["calls"], "share_of_step": top["ms"] / sum(r["ms"] for r in report),
            "peak_source": pk["source"]}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_acoustic_sample(steps, warmup):
    """Bounded sample of cfg2: B=2, Tx<=32 (cost is linear in padded frames x 100 steps)."""
    from oracle import oracle
    from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
    from promptttspp_b200.utils.synthetic import build_acoustic, synthetic_state_dict

    torch.set_num_threads(os.cpu_count() or 1)
    model = build_acoustic(bert=FixedPromptEmbedding(torch.zeros(1, 768)))
    sd = synthetic_state_dict(model, seed=1234)
    phoneme, lengths, cls_emb = cfg2_inputs(seed=2, B=2, lo=24, hi=33)
    g = torch.Generator().manual_seed(7)
    z_style = torch.randn(2, 1, 256, generator=g)
    cfg = dict(oracle.ACOUSTIC_CFG)
    times, frames = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        mel, _, _, flen = oracle.acoustic_infer_batch(sd, cfg, phoneme, lengths, cls_emb, z_style,
                                                      noise_fn=lambda s: torch.randn(s))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        frames = float(flen.sum())
        padded = mel.shape[0] * mel.shape[-1]
    t = statistics.median(times)
    return dict(value=frames / t, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"oracle/oracle.py acoustic_infer_batch, B=2 Tx<=32 -> {int(frames)} valid / {padded} padded "
                       f"frames, 100 diffusion steps, median of {len(times)} ({t:.2f} s each)"), t


def cpu_bigvgan_sample(steps, warmup):
    from oracle import oracle
    from promptttspp_b200.utils.synthetic import build_vocoder, synthetic_state_dict

    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic_state_dict(build_vocoder(), seed=4321)
    mel = cfg3_inputs(B=2, T=256)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.bigvgan_forward(sd, oracle.VOCODER_CFG, mel)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = statistics.median(times)
    audio_s = mel.shape[0] * mel.shape[-1] / 100.0
    return dict(rtf=t / audio_s, frames_per_sec=mel.shape[0] * mel.shape[-1] / t, cores=torch.get_num_threads(),
                kind="port", sample=f"oracle/oracle.py bigvgan_forward, B=2 x 256 frames ({audio_s:.2f} s audio), "
                                    f"median of {len(times)} ({t:.2f} s each)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    base, t = cpu_acoustic_sample(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "K_step": 100, "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bigvgan": cpu_bigvgan_sample(max(1, min(args.steps, 3)), 1),
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------

def run_native(args, rank, local_rank, world):
    import torch.distributed as dist

    from promptttspp_b200 import _abi

    torch.set_grad_enabled(False)
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _abi.lib()
    _abi.check(lib.pttspp_device_check())
    pk = peaks()
    model, voc = build_models(device)
    B = args.batch
    phoneme_h, lengths_h, cls_h = cfg2_inputs(seed=2 + rank, B=B)
    phoneme_h, lengths_h, cls_h = phoneme_h.pin_memory(), lengths_h.pin_memory(), cls_h.pin_memory()
    phoneme_d, lengths_d, cls_d = phoneme_h.to(device), lengths_h.to(device), cls_h.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        torch.manual_seed(1000 + rank)
        return model.infer_batch(phoneme_d, lengths_d, style_prompt=cls_d, use_max=True, noise_scale=0.5,
                                 return_f0=True)

    def step_e2e():
        torch.manual_seed(1000 + rank)
        p = phoneme_h.to(device, non_blocking=True)
        l = lengths_h.to(device, non_blocking=True)
        c = cls_h.to(device, non_blocking=True)
        mel, cf0, vuv, flen = model.infer_batch(p, l, style_prompt=c, use_max=True, noise_scale=0.5, return_f0=True)
        return mel.cpu(), flen.cpu()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            out = fn()
        barrier()
        lib.pttspp_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for _ in range(steps):
                out = fn()
            e1.record()
            barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out, lib.pttspp_launch_count(), clk.summary()

    if args.leg != "both":
        # profiling aid: exactly `warmup` + `steps` passes of one leg, nothing else launched (ncu -s/-c friendly)
        if args.leg == "acoustic":
            ms, out, launches, _ = timed(step_device, args.steps, args.warmup)
        else:
            mel_d = cfg3_inputs(seed=3 + rank).to(device)
            ms, out, launches, _ = timed(lambda: voc(mel_d), args.steps, args.warmup)
        if rank == 0:
            print(json.dumps({"leg": args.leg, "ms_per_step": ms, "gpu_launches_per_step": launches // args.steps}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- acoustic leg (headline) ----
    ms_dev, out, launches, clocks = timed(step_device, args.steps, args.warmup)
    flen = out[3]
    valid = torch.tensor([float(flen.sum())], device=device)
    padded = torch.tensor([float(out[0].shape[0] * out[0].shape[-1])], device=device)
    if world > 1:
        dist.all_reduce(valid)
        dist.all_reduce(padded)
    ms_e2e, out_e2e, _, _ = timed(step_e2e, args.steps, 1)
    h2d = phoneme_h.numel() * 8 + lengths_h.numel() * 8 + cls_h.numel() * 4
    d2h = out_e2e[0].numel() * 4 + out_e2e[1].numel() * 4
    # one extra, profiled step: per-family CUDA-event times + algorithmic work (not part of the timing above)
    lib.pttspp_prof_enable(1)
    step_device()
    rep_ac = prof_report(lib)
    lib.pttspp_prof_enable(0)

    # ---- BigVGAN leg ----
    mel_d = cfg3_inputs(seed=3 + rank).to(device)
    mel_hp = cfg3_inputs(seed=3 + rank).pin_memory()
    ms_voc, _, launches_voc, _ = timed(lambda: voc(mel_d), max(args.steps, 5), args.warmup)
    ms_voc_e2e, _, _, _ = timed(lambda: voc(mel_hp.to(device, non_blocking=True)).cpu(), max(args.steps, 5), 1)
    lib.pttspp_prof_enable(1)
    voc(mel_d)
    rep_voc = prof_report(lib)
    lib.pttspp_prof_enable(0)
    voc_frames = mel_d.shape[0] * mel_d.shape[-1] * world
    audio_s = voc_frames / 100.0

    if rank == 0:
        cpu_ac, _ = cpu_acoustic_sample(1, 0) if world == 1 and not args.no_cpu else (None, None)
        cpu_voc = cpu_bigvgan_sample(1, 0) if world == 1 and not args.no_cpu else None
        roof_ac = roofline_from(rep_ac, pk, in_long_step=True, rows=float(padded) / world)
        roof_voc = roofline_from(rep_voc, pk, in_long_step=True)
        frames_s = float(valid) / (ms_dev * 1e-3)
        line = {
            "metric": METRIC, "value": frames_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": B, "valid_frames": float(valid), "padded_frames": float(padded),
                       "K_step": 100, "l2": "working set (>1 GB of activations per step) exceeds the 126 MB L2",
                       "parallelism": f"utterance batches sharded over {world} GPU(s), no data-path collective"},
            "padded_frames_per_sec": float(padded) / (ms_dev * 1e-3),
            "diffnet_tflops_algorithmic": float(padded) * DIFFNET_FLOP_PER_FRAME / (ms_dev * 1e-3) / 1e12,
            "clocks": clocks,
            "e2e": {"value": float(valid) / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roof_ac,
            "kernel_families": rep_ac,
            "cpu_baseline": cpu_ac,
            "bigvgan": {
                "workload": "cfg3: BigVGAN 24 kHz vocoder only, batch 16 x 1024-frame mel",
                "rtf": (ms_voc * 1e-3) / (audio_s / world), "frames_per_sec": voc_frames / (ms_voc * 1e-3),
                "ms_per_step": ms_voc, "audio_seconds_per_step": audio_s,
                "tflops_algorithmic": voc_frames * BIGVGAN_FLOP_PER_FRAME / (ms_voc * 1e-3) / 1e12,
                "e2e": {"rtf": (ms_voc_e2e * 1e-3) / (audio_s / world), "ms_per_step": ms_voc_e2e,
                        "h2d_bytes_per_step": mel_hp.numel() * 4, "d2h_bytes_per_step": mel_hp.shape[0] * 240 * mel_hp.shape[-1] * 4},
                "gpu_launches": launches_voc, "roofline": roof_voc, "kernel_families": rep_voc,
                "cpu_baseline": cpu_voc,
            },
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--leg", default="both", choices=["both", "acoustic", "bigvgan"],
                    help="profiling aid (ncu launch lists): run only one leg; the default line carries both")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
