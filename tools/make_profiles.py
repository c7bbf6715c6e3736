"""Turn the artefacts of tools/gpu_round.sh (gpurun_out/) into the committed evidence under profiles/.

    python tools/make_profiles.py <tag>        # e.g. r01_final
"""
import csv
import gzip
import io
import json
import shutil
import subprocess
import sys
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import ncu_summary  # noqa: E402
import summarize_launches  # noqa: E402

OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"


def capture(fn, *a):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def main(tag):
    for name in ("bench.jsonl", "bench_ref.jsonl", "pytest_gpu.log"):
        if (OUT / name).exists():
            shutil.copy(OUT / name, PROF / f"{tag}_{name}")
    for leg in ("acoustic", "bigvgan"):
        src = OUT / f"launches_{leg}.csv"
        if src.exists():
            (PROF / f"{tag}_launches_{leg}_summary.txt").write_text(capture(summarize_launches.main, str(src)))
            with open(src, "rb") as f, gzip.open(PROF / f"{tag}_launches_{leg}.csv.gz", "wb") as g:
                g.write(f.read())
    traffic = {}
    for rep, out in (("full_umma_pair", "ncu_pair_kernels"), ("full_aa", "ncu_aa_snake")):
        if not (OUT / f"{rep}.ncu-rep").exists():
            continue
        raw = OUT / f"{rep}_raw.csv"
        with open(raw, "w") as f:
            subprocess.run(["ncu", "-i", str(OUT / f"{rep}.ncu-rep"), "--page", "raw", "--csv"], stdout=f, check=True)
        (PROF / f"{tag}_{out}.txt").write_text(capture(ncu_summary.main, str(raw)))
        rows = list(csv.reader(open(raw)))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            # tools/prof_umma.py launches the dilated+gate conv first, then the dual 1x1 projection
            key = ("aa_snake" if "aa_snake" in name else
                   "pair_dilated_gate" if "pair_dilated_gate" not in traffic else "pair_dual_1x1")
            traffic.setdefault(key, {"dram_bytes_per_launch": b, "duration_us": float(r[hdr.index("gpu__time_duration.sum")]),
                                     "kernel": name[:90]})
    if traffic:
        traffic["_note"] = ("ncu --set full --clock-control none on tools/prof_umma.py (DiffNet layer launches on 16 x 2048 = "
                            "32768 rows) and on the BigVGAN leg; dram__bytes_read.sum + dram__bytes_write.sum per launch")
        traffic["_rows"] = 32768
        (PROF / "r01_ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
    print("profiles written for", tag)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01_final")
