"""Print headline metrics + per-instruction stall hot spots of an .ncu-rep (needs -lineinfo build)."""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct", "lts__throughput.max.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct", "launch__registers_per_thread ",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ", "lts__t_sectors.sum ", "long_scoreboard_per_issue", "launch__grid_size", "launch__occupancy_limit",
        "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum ", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum ", "lts__t_sectors_srcunit_tex_op_red.sum ", "lts__d_sectors.max.pct", "lts__t_sector_throughput_srcunit_tex.avg"]
for h, u, v in zip(rows[0], rows[1], rows[-1]):
    if any(w in h + " " for w in want):
        print("%-90s %-10s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc, isamp, ils, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("stall_long_sb"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[isamp]), int(r[ils] or 0), r[isrc][:100], r[iex]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot)
for n, (s, l, sc, ex) in enumerate(data):
    if s > tot * thr:
        print("%4d %6d (%4.1f%%) long_sb %6d  exec %9s  %s" % (n, s, 100.0 * s / tot, l, ex, sc))
