"""Turn gpurun_out ncu artefacts into the committed summaries under profiles/.
usage: python tools/summarize_ncu.py <tag> (e.g. r01)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
for f in sorted(os.listdir(G)):
    if f.startswith(tag) and f.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", os.path.join(G, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        with open(os.path.join(P, f.replace(".ncu-rep", "_raw.txt")), "w") as fh:
            fh.write("# selected raw metrics of gpurun_out/%s (ncu --set full --clock-control none)\n" % f)
            for vals in rows[2:]:
                d = dict(zip(hdr, vals))
                fh.write("kernel: %s\n" % d.get("Kernel Name", "?"))
                for h, u, v in zip(hdr, units, vals):
                    if h in KEEP or "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                        fh.write("  %-80s %-16s %s\n" % (h, u, v))
        print("wrote", f)
for f in sorted(os.listdir(G)):
    if f.startswith(tag) and f.endswith(".csv"):
        txt = open(os.path.join(G, f)).read()
        open(os.path.join(P, f), "w").write(txt)
        print("copied", f)
