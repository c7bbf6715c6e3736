"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of one
BigVGAN forward (bench.py --leg bigvgan) per kernel family and merge it into profiles/r02_ncu_traffic.json
(`bigvgan_families`: what bench.py reports as bigvgan.roofline.traffic).

    python tools/bigvgan_traffic.py gpurun_out/bigvgan_traffic.csv
"""
import csv
import io
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}


def family(name):
    if "aa_snake" in name:
        return "aa_snake"
    if "umma" in name or "aa_conv" in name:
        return "conv1d_tcgen05_splitfp16"
    if "simt" in name:
        return "conv1d_simt_fp32"
    return "other"


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    per = defaultdict(lambda: defaultdict(float))  # launch id -> metric -> value
    names = {}
    for r in csv.DictReader(io.StringIO("".join(lines))):
        v = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"], 1.0)
        per[r["ID"]][r["Metric Name"]] = v
        names[r["ID"]] = r["Kernel Name"]
    fam = defaultdict(lambda: dict(launches=0, dram_bytes=0.0, ns=0.0))
    for i, m in per.items():
        f = fam[family(names[i])]
        f["launches"] += 1
        f["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        f["ns"] += m.get("gpu__time_duration.sum", 0.0)
    out = {}
    for k, f in fam.items():
        out[k] = dict(launches=f["launches"], dram_bytes_total=f["dram_bytes"],
                      dram_bytes_per_launch=f["dram_bytes"] / f["launches"], duration_us_total=f["ns"] / 1e3,
                      dram_gbps=f["dram_bytes"] / max(f["ns"], 1.0))
        print(f"{k:28s} {f['launches']:4d} launches  {f['dram_bytes'] / 1e9:8.2f} GB  {f['ns'] / 1e6:8.2f} ms  "
              f"{f['dram_bytes'] / max(f['ns'], 1.0):7.1f} GB/s")
    # algorithmic bytes of cfg3 (B=16 x 1024 frames): per AMP layer conv1 reads 4 + writes 4, conv2 reads 4 + residual 4 +
    # writes 4 bytes per element (the operand planes are 2 x 2 bytes), activations 4 + 4
    elems = [16 * 6144 * 256, 16 * 30720 * 128, 16 * 122880 * 64, 16 * 245760 * 32]
    conv_alg = sum(e * 9 * 20 for e in elems)
    aa_alg = sum(e * 18 * 8 for e in elems) + elems[-1] * 8
    if "conv1d_tcgen05_splitfp16" in out:
        out["conv1d_tcgen05_splitfp16"]["algorithmic_bytes_per_launch"] = conv_alg / out["conv1d_tcgen05_splitfp16"]["launches"]
        out["conv1d_tcgen05_splitfp16"]["algorithmic_bytes_total"] = conv_alg
    if "aa_snake" in out:
        out["aa_snake"]["algorithmic_bytes_per_launch"] = aa_alg / out["aa_snake"]["launches"]
        out["aa_snake"]["algorithmic_bytes_total"] = aa_alg
    p = ROOT / "profiles" / "r02_ncu_traffic.json"
    t = json.loads(p.read_text())
    t["bigvgan_families"] = out
    t["_note_bigvgan"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                          "over the 165 launches of one cfg3 BigVGAN forward (bench.py --leg bigvgan), per kernel family; "
                          "the AMP-layer convs only in the algorithmic figure (transposed convs, conv_pre/post excluded)")
    p.write_text(json.dumps(t, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
