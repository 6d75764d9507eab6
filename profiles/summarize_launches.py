"""Per-step kernel time table from an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --no-graph`.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [step_index] [--each]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 2:] if len(r) > vi]
    names = [re.sub(r"^void |d3b::|\(anonymous namespace\)::", "", re.sub(r"\(.*", "", k)) for k, _v in data]
    return names, [v for _k, v in data]


def main(path, step=0, each=False):
    names, times = load(path)
    starts = [i for i, n in enumerate(names) if "vox_insert" in n]
    s, e = starts[step], starts[step + 1]
    agg = collections.OrderedDict()
    for n, v in zip(names[s:e], times[s:e]):
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += v
        agg[n][1] += 1
    tot = sum(v[0] for v in agg.values())
    print("| kernel | launches | time [us] | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("| %s | %d | %.1f | %.1f %% |" % (k[:80], v[1], v[0] / 1000, 100 * v[0] / tot))
    print("| **total** | %d | %.1f | |" % (e - s, tot / 1000))
    if each:
        print("\nconvolution launches in order [us]:", [round(v / 1000, 1) for n, v in zip(names[s:e], times[s:e])
                                                       if "spconv" in n or "bev_conv" in n])


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0], int(args[1]) if len(args) > 1 else 0, "--each" in sys.argv)
