"""profiles/traffic.json from an ncu launch list of one decode step.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/prof_decode.py
    python tools/traffic_from_ncu.py gpurun_out/launches.csv profiles/r02_launches.csv

Writes, per kernel class of bench.py, the mean DRAM bytes (read + write) per launch and the launch count, next to the hash of
the library sources the capture was taken with (bench.py refuses nothing: it prints the note so a reader can tell whether the
capture matches the build).  Also prints the per-kernel summary that goes to profiles/*_summary.txt."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = {"actconv": ("conv_umma_fused_kernel", "actconv_tc_kernel"), "act1d": ("act1d_c8t_kernel", "act1d_tc_kernel", "act1d_kernel"),
           "conv1d": ("conv_umma_kernel",)}


def main():
    src = sys.argv[1]
    rows = list(csv.reader(open(src, errors="replace")))
    hdr = next(r for r in rows if "Kernel Name" in r)
    i_id, i_name, i_metric, i_val, i_unit = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = {}
    for r in rows:
        if len(r) != len(hdr) or r is hdr or r[i_id] == "ID":
            continue
        d = launches.setdefault(r[i_id], {"name": r[i_name]})
        v = float(r[i_val].replace(",", ""))
        unit = r[i_unit].lower()
        if "byte" in unit:                                        # ncu scales units in the csv (Kbyte / Mbyte / Gbyte)
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        if unit in ("us", "usecond"):
            v *= 1e3
        if unit in ("ms", "msecond"):
            v *= 1e6
        d[r[i_metric]] = v
    per = {}
    for d in launches.values():
        short = d["name"].split("(")[0].split("::")[-1]
        e = per.setdefault(short, {"n": 0, "ns": 0.0, "bytes": 0.0})
        e["n"] += 1
        e["ns"] += d.get("gpu__time_duration.sum", 0.0)
        e["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total_ns = sum(e["ns"] for e in per.values()) or 1.0
    for k, e in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"{k:44s} n={e['n']:4d} total={e['ns'] / 1e6:9.3f} ms share={100 * e['ns'] / total_ns:5.1f}%  "
              f"dram={e['bytes'] / 1e9:8.3f} GB  ({e['bytes'] / max(e['ns'], 1):7.1f} GB/s under ncu)")
    print(f"total {total_ns / 1e6:.3f} ms over {sum(e['n'] for e in per.values())} launches")
    out = {}
    for cls, names in CLASSES.items():
        sel = [e for k, e in per.items() if any(k.startswith(n) for n in names)]
        n = sum(e["n"] for e in sel)
        if n:
            out[cls] = {"bytes_per_launch": sum(e["bytes"] for e in sel) / n, "launches": n,
                        "source": os.path.relpath(sys.argv[2] if len(sys.argv) > 2 else src, ROOT)}
    h = os.path.join(ROOT, "index-tts-ipex_b200", "lib", "libbigvgan_b200.so.srchash")
    out["library"] = open(h).read().strip()[:16] if os.path.exists(h) else "?"
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("wrote profiles/traffic.json:", json.dumps(out))


if __name__ == "__main__":
    main()
