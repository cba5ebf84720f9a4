"""python tools/ncu_summary.py <report.ncu-rep> [...] - the handful of metrics DESIGN.md / bench.py quote, per launch, as text.
Runs `ncu -i ... --page raw --csv` (no GPU needed) and prints one block per kernel launch."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 LSU data-pipe wavefronts %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts for tensor operands %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "UTCHMMA tf32 ops % of peak"),
    ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "UTCHMMA fp16 ops % of peak"),
    ("smsp__sass_inst_executed_op_utcmma.sum", "tcgen05.mma instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (warps/issue)"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        print("== %s (%d launch(es); ncu --set full --clock-control none; cold-cache, serialised - compare shares, not absolutes)" % (rep.split("/")[-1], len(data)))
        for r in data:
            print("-- %s" % r[ix["Kernel Name"]][:110])
            for k, label in KEYS:
                if k in ix and r[ix[k]] not in ("", "0", "n/a"):
                    print("   %-46s %s %s" % (label, r[ix[k]], units[ix[k]]))


if __name__ == "__main__":
    main()
