"""Key metrics of every kernel in an `ncu --set full` report, as text for profiles/.

    python tools/ncu_keymetrics.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/x_ncu_full.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM read bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_op_ldgsts.sum", "LDGSTS (cp.async) instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per instruction issued (cycles)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"== {rep}: no kernels")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"== {rep}")
        for r in rows[2:]:
            print(f"-- {r[col['Kernel Name']][:150]}")
            for k, label in KEYS:
                if k in col and r[col[k]] != "":
                    print(f"   {label:48s} {r[col[k]]} {units[col[k]]}")
            if "dram__bytes_read.sum" in col and "gpu__time_duration.sum" in col:
                def val(k):
                    v, u = float(r[col[k]].replace(",", "")), units[col[k]]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
                                "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}.get(u, 1)
                t = val("gpu__time_duration.sum")
                b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                print(f"   {'DRAM bytes / duration':48s} {b / t / 1e9:.1f} GB/s")


if __name__ == "__main__":
    main()
