"""Text summary of one .ncu-rep capture (the metrics the notes and DESIGN.md quote): python tools/ncu_summary.py REP "header line" > profiles/rNN_ncu_full_<what>_summary.txt"""
import csv
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__shared_mem_config_size", "lts__throughput", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__icc_request_hit_rate", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "sm__throughput", "sm__warps_active", "smsp__inst_executed.sum",
        "smsp__issue_active", "smsp__average_warps_issue_stalled", "sm__inst_executed.avg.per_cycle", "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warps_eligible")


def main():
    rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    if header:
        print("# " + header)
    d = dict(zip(names, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", ""))[1])
    for k in sorted(d):
        if k.startswith(KEEP) and "per_issue_active" not in k or ("smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")):
            u, v = d[k]
            print(f"{k} = {v} {u}".rstrip())


if __name__ == "__main__":
    main()
