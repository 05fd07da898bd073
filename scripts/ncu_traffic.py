"""Extract per-launch DRAM traffic and pipe utilisation from an `ncu --set full` capture of scripts/ncu_kernels.py and write
profiles/rNN_ncu_traffic.json (read by bench.py for `roofline.traffic`) plus a markdown table.

    python scripts/ncu_traffic.py gpurun_out/full_kernels.ncu-rep profiles/r02_ncu_traffic.json [profiles/r02_ncu_full.md]

Runs on the CPU box (ncu -i only reads the report).  The JSON records the sha256 of world-in-world_b200/csrc/wiw_gemm.cu at
extraction time, which must be the source the capture was built from (extract right after the gpurun call, before editing the
kernel again): bench.py refuses a JSON whose hash differs from the current source.
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# launches of scripts/ncu_kernels.py in order (each shape twice: warm-up, then the one that is listed)
LABELS = [
    ("conv3x3 28x72x128 320->320 (K=2880, N=320)", 2.0 * 258048 * 320 * 2880, 2 * 258048 * 320 * 2 + 320 * 2880 * 2),
    ("linear qkv M=258048 K=320 N=960", 2.0 * 258048 * 960 * 320, 258048 * (320 + 960) * 2 + 960 * 320 * 2),
    ("linear ff1+GEGLU M=258048 K=320 N=2560", 2.0 * 258048 * 2560 * 320, 258048 * (320 + 1280) * 2 + 2560 * 320 * 2),
    ("linear ff2+residual M=258048 K=1280 N=320", 2.0 * 258048 * 320 * 1280, 258048 * (1280 + 320 + 320) * 2 + 320 * 1280 * 2),
    ("attn_spatial 4 img x 5 heads, S=9216, d=64", 4.0 * 4 * 5 * 9216 * 9216 * 64, 4 * 9216 * (960 + 320) * 2),
]
METRICS = {
    "gpu__time_duration.sum": "ns",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active": "tensor_pct_alt",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "mufu_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "launch__registers_per_thread": "regs",
    "sm__cycles_elapsed.max": "cycles",
}


def to_bytes(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out_json = sys.argv[1], sys.argv[2]
    out_md = sys.argv[3] if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in data:
        d = {"kernel": r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), "grid": r[ix["Grid Size"]] if "Grid Size" in ix else ""}
        for m, k in METRICS.items():
            if m in ix and r[ix[m]] not in ("", "n/a", "no data"):
                d[k] = to_bytes(r[ix[m]], units[ix[m]]) if k in ("rd", "wr") else float(r[ix[m]].replace(",", ""))
                if k == "ns":
                    d[k] *= {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(units[ix[m]], 1)
        launches.append(d)
    listed = launches[1::2]                                 # the second launch of every shape
    table = []
    for (label, flop, abytes), d in zip(LABELS, listed):
        us = d["ns"] / 1e3
        table.append({"launch": label, "kernel": d["kernel"], "time_us": round(us, 1), "tflops": round(flop / us / 1e6, 0),
                      "dram_read_bytes": d.get("rd"), "dram_write_bytes": d.get("wr"), "dram_bytes": (d.get("rd", 0) + d.get("wr", 0)),
                      "algorithmic_bytes": abytes, "tensor_pipe_active_pct": d.get("tensor_pct", d.get("tensor_pct_alt")),
                      "issue_active_pct": d.get("issue_pct"), "mufu_pct": d.get("mufu_pct"), "dram_pct": d.get("dram_pct"),
                      "l2_pct": d.get("l2_pct"), "regs": d.get("regs")})
    sha = hashlib.sha256(open(os.path.join(REPO, "world-in-world_b200", "csrc", "wiw_gemm.cu"), "rb").read()).hexdigest()[:16]
    doc = {"source_report": os.path.basename(rep), "kernel_source_sha16": sha, "dominant": table[0], "launches": table,
           "how": "ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|attn_spatial' python scripts/ncu_kernels.py all; "
                  "second launch of each shape; bytes are dram__bytes_read.sum / dram__bytes_write.sum per launch"}
    json.dump(doc, open(out_json, "w"), indent=1)
    if out_md:
        with open(out_md, "w") as f:
            f.write(f"# ncu --set full, {os.path.basename(rep)} (kernel source {sha})\n\n")
            f.write("| launch | time us | TFLOP/s | tensor pipe % | issue % | MUFU % | DRAM read MB | DRAM write MB | algorithmic MB | DRAM % | L2 % | regs |\n")
            f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
            for t in table:
                f.write(f"| {t['launch']} | {t['time_us']} | {t['tflops']:.0f} | {t['tensor_pipe_active_pct']} | {t['issue_active_pct']} | {t['mufu_pct']} | "
                        f"{(t['dram_read_bytes'] or 0) / 1e6:.1f} | {(t['dram_write_bytes'] or 0) / 1e6:.1f} | {t['algorithmic_bytes'] / 1e6:.1f} | "
                        f"{t['dram_pct']} | {t['l2_pct']} | {t['regs']} |\n")
    print(json.dumps(doc["dominant"]))


if __name__ == "__main__":
    main()
