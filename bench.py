"""Benchmark of the PromptTTS++ inference hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Headline workload (N=1): BASELINE.json configs[1] -- prompttts_mdn_v2 acoustic model, batch 16 synthetic
phoneme sequences (len <= 256), full text -> mel forward incl. the 100-step diffusion decoder; metric =
valid mel frames per second.  The same JSON line carries the BigVGAN leg of the metric (configs[2]:
16 x 1024-frame mel -> 24 kHz waveform, real-time factor) under "bigvgan".
A "step" is one pass of the path over one batch.  N > 1 (torchrun, one rank per GPU): every rank runs its
own batch (weak scaling, utterances are independent -- no data-path collective); time = max over ranks.
The line also carries BASELINE.json's other configs: "cfg1" (one 50-phoneme utterance: latency), "cfg4" (32 variable-length
utterances, text ids -> waveform through serving.BatchedSynthesizer, plain and F0-aware vocoder) and "cfg5" (256
utterances through dist.synthesize_sharded over the N ranks: strong scaling).
`--impl reference` times the UNMODIFIED reference modules installed in baseline/_ref (baseline/reference_arm.py) on the
host cores; the oracle port is only the fallback when that install is missing.
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
TAGS = ["conv1d_simt_fp32", "conv1d_tcgen05_splitfp16", "aa_snake", "layernorm", "relpos_attention", "other",
        "diffnet_layers_tcgen05_splitfp16"]
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
    n = len(TAGS)
    ms, fl, by = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
    calls = (ctypes.c_int64 * n)()
    from promptttspp_b200 import _abi

    _abi.check(lib.pttspp_prof_report(ms, fl, by, calls, n))
    return [dict(kernel=TAGS[i], ms=ms[i], flops=fl[i], bytes=by[i], calls=calls[i]) for i in range(n)]


def ncu_traffic(kernel, rows, leg="acoustic"):
    """DRAM bytes per launch of `kernel`'s family from the committed ncu --set full capture (profiles/), scaled from
    the capture's row count to this run's; None when no capture covers the family."""
    p2 = ROOT / "profiles" / "r02_ncu_traffic.json"
    if leg == "bigvgan" and p2.exists():
        fam = json.loads(p2.read_text()).get("bigvgan_families", {}).get(kernel)
        if fam:
            return fam["dram_bytes_per_launch"], (
                f"profiles/r02_ncu_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum over the {fam['launches']} "
                f"launches of this family in one cfg3 forward, mean per launch; algorithmic bytes per launch "
                f"{fam.get('algorithmic_bytes_per_launch')})")
        return None, None
    if kernel == TAGS[6] and p2.exists() and rows:
        t = json.loads(p2.read_text())["diffnet_layers_kernel"]
        return t["dram_bytes_per_launch"] * rows / t["rows"], (
            f"profiles/r02_ncu_traffic.json (ncu --set full capture of one diffnet_layers_kernel launch at {t['rows']} "
            f"rows: dram__bytes_read.sum + dram__bytes_write.sum), scaled to {int(rows)} rows")
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


def roofline_from(report, pk, in_long_step, rows=None, leg="acoustic"):
    """Dominant kernel family of one profiled step -> the `roofline` object."""
    top = max(report, key=lambda r: r["ms"])
    if top["ms"] <= 0 or top["calls"] == 0:
        return None
    traffic, traffic_src = ncu_traffic(top["kernel"], rows, leg)
    if top["flops"] > 0:
        peak = pk["tf_sustained"] if in_long_step else pk["tf_burst"]
        ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": top["kernel"], "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "launches": top["calls"],
                "avg_launch_ms": top["ms"] / top["calls"], "share_of_step": top["ms"] / sum(r["ms"] for r in report),
                "mma_frac": (3.0 * ach / peak) if top["kernel"] in (TAGS[1], TAGS[6]) else None,
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
    from baseline import reference_arm

    return reference_arm.acoustic_sample(steps, warmup)


def cpu_bigvgan_sample(steps, warmup):
    from baseline import reference_arm

    return reference_arm.bigvgan_sample(steps, warmup)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from baseline import reference_arm

    # each step is a bounded sample (>= 2 k padded frames, ~10 s of CPU work): at most 1 warm-up + 4 timed repetitions
    # so that the arm ends within a few minutes whatever --steps / --warmup the native arm was given
    steps, warmup = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    base, t = reference_arm.acoustic_sample(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "K_step": 100, "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bigvgan": reference_arm.bigvgan_sample(3, 1),
        "cfg1": reference_arm.cfg1_sample(3, 1),
    }
    print(json.dumps(line))



# ---------------------------------------------------------------------------------------------
# BASELINE.json's other configs (reported inside the headline line)
# ---------------------------------------------------------------------------------------------

def var_len_requests(n, seed, lo=32, hi=257):
    """n synthetic requests: phoneme id sequences of length [lo, hi) + sentence embeddings (host tensors)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi, (n,), generator=g).tolist()
    return [torch.randint(3, 90, (k,), generator=g) for k in lens], torch.randn(n, 768, generator=g)


def _event_time(fn, device):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1), out


def leg_cfg1(model, device, reps=7):
    """configs[0]: ONE utterance of 50 phonemes through `infer` (the app.py:56-82 call): latency."""
    from baseline import reference_arm

    phoneme, _, cls = reference_arm._inputs(11, 1, 50, 51)
    ph_h, cls_h = phoneme.pin_memory(), cls.pin_memory()
    ph_d, cls_d = phoneme.to(device), cls.to(device)

    def dev():
        torch.manual_seed(7)
        return model.infer(ph_d, style_prompt=cls_d, use_max=True, noise_scale=0.5)

    def e2e():
        torch.manual_seed(7)
        return model.infer(ph_h.to(device, non_blocking=True), style_prompt=cls_h.to(device, non_blocking=True),
                           use_max=True, noise_scale=0.5).cpu()

    for _ in range(2):
        dev()
    t_dev = sorted(_event_time(dev, device)[0] for _ in range(reps))
    t_e2e = sorted(_event_time(e2e, device)[0] for _ in range(reps))
    frames = int(dev().shape[-1])
    ms, ms_e = t_dev[len(t_dev) // 2], t_e2e[len(t_e2e) // 2]
    return {"workload": "cfg1: single utterance, 50 synthetic phonemes, fixed style-prompt embedding, acoustic -> mel "
                        "(model.infer, 100 diffusion steps)",
            "latency_ms": ms, "latency_ms_min": t_dev[0], "frames": frames, "frames_per_sec": frames / (ms * 1e-3),
            "e2e": {"latency_ms": ms_e, "frames_per_sec": frames / (ms_e * 1e-3), "h2d_bytes": 50 * 8 + 768 * 4,
                    "d2h_bytes": frames * 80 * 4},
            "reps": reps}


def leg_cfg4(model, voc, voc_f0, device, n=32):
    """configs[3]: text ids + style-prompt embeddings -> waveforms for 32 variable-length requests through
    serving.BatchedSynthesizer (token-bucketed batches, acoustic -> f0 post-processing -> vocoder, one D2H per batch).
    Host tensors in, host waveforms out: the timed region holds every copy."""
    from promptttspp_b200.serving import BatchedSynthesizer, MelStats

    phonemes, emb = var_len_requests(n, seed=41)
    res = {"workload": f"cfg4: end-to-end text+style-prompt -> 24 kHz waveform, {n} variable-length requests "
                       "(32..256 phonemes) through serving.BatchedSynthesizer (max_tokens 8192, batch boundaries "
                       "minimising the padded work: serving.cost_buckets)"}
    for name, v in (("bigvgan", voc), ("bigvgan_f0", voc_f0)):
        srv = BatchedSynthesizer(model, v, MelStats(mean=-5.0, std=2.0), max_tokens=8192, max_sentences=32)
        res["batches"] = [len(b) for b in __import__("promptttspp_b200.serving", fromlist=["cost_buckets"]).cost_buckets(
            [int(p.numel()) for p in phonemes], 8192, 32)]

        def run():
            torch.manual_seed(11)
            return srv.synthesize(phonemes, emb, device=device)

        run()
        ms, out = _event_time(run, device)
        samples = sum(int(w.numel()) for w in out.values())
        frames = samples // 240
        res[name] = {"ms": ms, "valid_frames": frames, "frames_per_sec": frames / (ms * 1e-3),
                     "audio_seconds": samples / 24000.0, "rtf": (ms * 1e-3) / (samples / 24000.0),
                     "h2d_bytes": sum(p.numel() for p in phonemes) * 8 + emb.numel() * 4, "d2h_bytes": samples * 4}
    return res


def leg_cfg5(model, voc, device, rank, world, n=256, batch_size=16):
    """configs[4]: 256 utterances sharded over the `world` ranks with dist.synthesize_sharded (length-sorted round-robin
    deal, batches of 16 neighbouring lengths, no data-path collective), text ids -> waveform; STRONG scaling: the job
    is the same 256 utterances at every N, time = max over ranks."""
    import torch.distributed as dist

    from promptttspp_b200.dist import gather_frame_counts, synthesize_sharded

    phonemes, emb = var_len_requests(n, seed=51, lo=64)
    emb_d = emb.to(device)

    def synth_batch(padded, lens, idx):
        torch.manual_seed(1000 + idx[0])  # the draw depends on the batch, not on the rank that runs it
        p = padded.pin_memory().to(device, non_blocking=True)
        l = lens.pin_memory().to(device, non_blocking=True)
        mel, flen = model.infer_batch(p, l, style_prompt=emb_d[idx], use_max=True, noise_scale=0.5)
        wav = voc(mel * 2.0 - 5.0).squeeze(1).cpu()
        nf = flen.cpu().long().tolist()
        return [(nf[b], float(wav[b, : nf[b] * 240].double().abs().sum())) for b in range(len(idx))]

    synthesize_sharded(phonemes[: 2 * world], synth_batch, batch_size, world, rank)  # warm-up: one small batch
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mine = synthesize_sharded(phonemes, synth_batch, batch_size, world, rank)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    counts = gather_frame_counts({i: r[0] for i, r in mine.items()}, n)
    frames = sum(counts)
    return {"workload": f"cfg5: {n} utterances (64..256 phonemes), text ids -> waveform, dist.synthesize_sharded over "
                        f"{world} rank(s), batches of {batch_size}",
            "scaling": "strong", "n_gpus": world, "ms": float(ms), "valid_frames": frames,
            "frames_per_sec": frames / (float(ms) * 1e-3), "audio_seconds": frames / 100.0,
            "rtf": float(ms) * 1e-3 / (frames / 100.0), "utterances_this_rank": len(mine),
            "frame_count_checksum": int(sum((i + 1) * c for i, c in enumerate(counts)) % 1000000007)}

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

    # ---- the other BASELINE configs ----
    extra = {}
    if not args.no_extra:
        from golden_cases import F0_KWARGS
        from promptttspp_b200.utils.synthetic import build_vocoder_f0, synthetic_state_dict

        if rank == 0:
            extra["cfg1"] = leg_cfg1(model, device)
            voc_f0 = build_vocoder_f0(**F0_KWARGS)
            voc_f0.load_state_dict(synthetic_state_dict(voc_f0, seed=4322), strict=True)
            extra["cfg4"] = leg_cfg4(model, voc, voc_f0.to(device).eval(), device)
            del voc_f0
        barrier()
        extra["cfg5"] = leg_cfg5(model, voc, device, rank, world)

    if rank == 0:
        cpu_ac, _ = cpu_acoustic_sample(2, 1) if world == 1 and not args.no_cpu else (None, None)
        cpu_voc = cpu_bigvgan_sample(2, 1) if world == 1 and not args.no_cpu else None
        if "cfg1" in extra and world == 1 and not args.no_cpu:
            from baseline import reference_arm

            extra["cfg1"]["cpu_baseline"] = reference_arm.cfg1_sample(2, 1)
        roof_ac = roofline_from(rep_ac, pk, in_long_step=True, rows=float(padded) / world)
        roof_voc = roofline_from(rep_voc, pk, in_long_step=True, leg="bigvgan")
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
        line.update(extra)
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
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg1 / cfg4 / cfg5 legs")
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
