"""Key raw metrics of the (first) kernel in an ncu report as one JSON object.  Usage: ncu_metrics.py <report>"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u, v = rows[0], rows[1], rows[2]
m = {a: (b, c) for a, b, c in zip(h, u, v)}
def val(k, scale=None):
    if k not in m:
        return None
    unit, x = m[k]
    try:
        x = float(x.replace(",", ""))
    except ValueError:
        return None
    if scale == "bytes":
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    if scale == "ms":
        x *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)
    return x
print(json.dumps({"kernel": m.get("Kernel Name", ("", ""))[1],
                  "ms": val("gpu__time_duration.sum", "ms"),
                  "dram_read": val("dram__bytes_read.sum", "bytes"), "dram_write": val("dram__bytes_write.sum", "bytes"),
                  "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "lanes_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
                  "warp_inst": val("smsp__inst_executed.sum"), "regs": val("launch__registers_per_thread"),
                  "grid": val("launch__grid_size"), "block": val("launch__block_size"),
                  "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "local_ld_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
                  "local_st_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum")}))
