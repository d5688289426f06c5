"""Print the metrics DESIGN.md / profiles cite from an `ncu --page raw --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]
for k in KEYS:
    hit = [h for h in d if h == k or h.endswith("." + k)]
    for h in hit[:1]:
        print("| `%s` | %s | %s |" % (k, d[h][1], d[h][0]))
