"""Turns the ncu reports of a gpurun session into the committed text summaries under profiles/.

    python tools/ncu_summary.py gpurun_out/r1_ncu_gemm_tc_v6.ncu-rep profiles/r1_ncu_gemm_tc_v6_summary.txt
    python tools/ncu_summary.py --launches gpurun_out/r1_ncu_launches_v6.csv profiles/r1_ncu_launches_v6_summary.txt
    python tools/ncu_summary.py --traffic gpurun_out/r1_ncu_gemv_tma_v6.ncu-rep profiles/r1_ncu_traffic.json

--traffic writes, for the captured launch with the most DRAM bytes, dram__bytes_read.sum + dram__bytes_write.sum: bench.py puts it in
roofline.traffic next to that launch's algorithmic bytes."""
import collections, csv, io, json, subprocess, sys

WANT = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dsmem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts")]


def raw(rep):
    # a report, or the CSV that `ncu -i <rep> --page raw --csv` printed on the GPU box (reports of 100+ launches do not travel)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    a = sys.argv[1:]
    if a[0] == "--launches":
        rows = [r for r in csv.reader(open(a[1])) if len(r) > 5]
        hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            e = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); e[0] += 1; e[1] += v
        tot = sum(e[1] for e in agg.values())
        lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none, CUDA graphs off; kernel | launches | total us | share of the run"]
        lines += [f"{nm} | {c} | {t / 1e3:.1f} | {t / tot:.3f}" for nm, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])]
        open(a[2], "w").write("\n".join(lines) + "\n")
    elif a[0] == "--traffic":
        hdr, units, rows = raw(a[1])
        rd, wr, nm, du = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
        best = max(rows, key=lambda r: to_bytes(r[rd], units[rd]))
        json.dump({"kernel": best[nm], "report": a[1].split("/")[-1], "dram_bytes_read": to_bytes(best[rd], units[rd]), "dram_bytes_write": to_bytes(best[wr], units[wr]),
                   "dram_bytes": to_bytes(best[rd], units[rd]) + to_bytes(best[wr], units[wr]), "duration_us_under_ncu": float(best[du]),
                   "grid": int(best[hdr.index("launch__grid_size")])}, open(a[2], "w"), indent=1)
    else:
        hdr, units, rows = raw(a[0])
        cols = [(k, n) for k, n in WANT if k in hdr]
        lines = ["# ncu --set full --clock-control none --import-source on, one row per captured launch; per-launch numbers are cold-cache and",
                 "# serialised (no programmatic-dependent-launch overlap): they explain shares, they are not bench values", "",
                 " | ".join(["kernel"] + [f"{n} [{units[hdr.index(k)]}]" for k, n in cols])]
        lines += [" | ".join([r[hdr.index("Kernel Name")][:40]] + [r[hdr.index(k)] for k, n in cols]) for r in rows]
        open(a[1], "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
