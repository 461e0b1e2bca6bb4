"""One-kernel summary of an `ncu --set full` report (read here, no GPU): duration, DRAM traffic, tensor /
issue activity, occupancy, top stall reasons.   python tools/ncu_summary.py rep.ncu-rep > profiles/x.md"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        def g(k):
            return d.get(k, ("", "n/a"))
        print("## `%s`\n" % d["Kernel Name"][1][:150])
        print("grid %s x block %s, %s registers/thread, dynamic smem %s %s\n" % (
            g("launch__grid_size")[1], g("launch__block_size")[1], g("launch__registers_per_thread")[1],
            g("launch__shared_mem_per_block_dynamic")[1], g("launch__shared_mem_per_block_dynamic")[0]))
        keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum"]
        print("| metric | value | unit |\n|---|---:|---|")
        for k in keys:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k][1], d[k][0]))
        stalls = [(h, float(v[1].replace(",", ""))) for h, v in d.items()
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[1] not in ("", "n/a")]
        stalls.sort(key=lambda x: -x[1])
        print("\ntop stall reasons (warps per issue-active cycle): " +
              ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                        for h, v in stalls[:5]))
        print()

if __name__ == "__main__":
    main()
