"""Per-step kernel time table from an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --no-graph`.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [step_index]"""
import collections
import csv
import re
import sys


def main(path, step=1):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 2:] if len(r) > vi]
    starts = [i for i, d in enumerate(data) if "vox_insert" in d[0]]
    st = data[starts[step]:starts[step + 1]]
    agg = collections.OrderedDict()
    for k, v in st:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void |d3b::|\(anonymous namespace\)::", "", k)
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += v
        agg[k][1] += 1
    tot = sum(v[0] for v in agg.values())
    print("| kernel | launches | time [us] | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("| %s | %d | %.1f | %.1f %% |" % (k[:80], v[1], v[0] / 1000, 100 * v[0] / tot))
    print("| **total** | %d | %.1f | |" % (len(st), tot / 1000))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
