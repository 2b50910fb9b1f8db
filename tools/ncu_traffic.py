"""DRAM traffic and headline metrics of one kernel from an ncu report -> profiles/*.json / *.txt
usage: python tools/ncu_traffic.py REPORT.ncu-rep OUT.json [label]      (reads with `ncu -i ... --page raw --csv`)"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
label = sys.argv[3] if len(sys.argv) > 3 else rep
txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, r = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
def num(k):
    v, u = d[k]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}.get(u, 1.0)
    return x * scale
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.sum", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
res = {"kernel": d.get("Kernel Name", ("?", ""))[0], "source": label}
for k in keep:
    if k in d:
        try:
            res[k] = num(k)
        except ValueError:
            res[k] = d[k][0]
res["dram_bytes_per_launch"] = res.get("dram__bytes_read.sum", 0.0) + res.get("dram__bytes_write.sum", 0.0)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
