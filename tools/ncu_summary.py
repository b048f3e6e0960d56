"""Summarise ncu captures into markdown for profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_summary.py full /tmp/raw.csv
    python tools/ncu_summary.py launches gpurun_out/launches.csv
"""
import collections
import csv
import re
import sys

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem")]


def short(name):
    name = re.sub(r"\bdrl::", "", name)
    name = re.sub(r"\(T2.*", "", name)
    return name.replace("void ", "")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(lbl for _, lbl in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in data:
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v, u = r[idx[k]], units[idx[k]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append((v + " " + u).strip())
            else:
                cells.append("n/a")
        print("| `%s` | %s |" % (short(r[idx["Kernel Name"]]), " | ".join(cells)))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(short(row["Kernel Name"])[:160], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, us, 100 * us / tot))


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
