"""ncu report -> small markdown summary (run where ncu is installed; no GPU needed).

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_xxx.md [title]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), n, units[hdr.index(m)]) for m, n in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as fh:
        fh.write("# %s\n\nsource: `ncu --set full --clock-control none` (per-launch, cold cache, serialised); units in header.\n\n" % title)
        fh.write("| kernel | " + " | ".join("%s [%s]" % (n, u) for _, n, u in cols) + " |\n")
        fh.write("|---|" + "---|" * len(cols) + "\n")
        for r in body:
            name = r[ki].split("(")[0].replace("void ", "").replace("d3b::", "")[:48]
            vals = []
            for i, _, _ in cols:
                try:
                    vals.append("%.4g" % float(r[i].replace(",", "")))
                except ValueError:
                    vals.append(r[i])
            fh.write("| %s | %s |\n" % (name, " | ".join(vals)))
    print("wrote", out, len(body), "launches")


if __name__ == "__main__":
    main()
