"""Summarise an .ncu-rep (raw page) into the handful of metrics we steer by (one JSON object per launch)."""
import csv, json, subprocess, sys
WANT = {
    'gpu__time_duration.sum': 'ms',
    'launch__registers_per_thread': 'regs',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__occupancy_limit_registers': 'occ_limit_regs',
    'launch__occupancy_limit_shared_mem': 'occ_limit_smem',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed': 'pipe_fmaheavy_pct',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed': 'pipe_alu_pct',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'inst_fma_pct',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'inst_alu_pct',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active': 'inst_fp64_pct',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'inst_lsu_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'smsp__inst_executed.sum': 'warp_insts',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'gpc__cycles_elapsed.avg.per_second': 'gpc_ghz',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_pipe',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_sb',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_sb',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'stall_wait',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'stall_mio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio': 'stall_not_selected',
}
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = {'kernel': vals[hdr.index('Kernel Name')][:90]}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT and v != '':
                try:
                    x = float(v.replace(',', ''))
                except ValueError:
                    continue
                key = WANT[h]
                if key.startswith('dram_') and key != 'dram_pct':
                    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(u, 1.0)
                    x *= mult
                    key += '_bytes'
                rec[key] = round(x, 4) if abs(x) < 1e6 else x
        if 'dram_read_bytes' in rec and 'dram_write_bytes' in rec:
            rec['dram_traffic_bytes'] = rec['dram_read_bytes'] + rec['dram_write_bytes']
            if 'ms' in rec:
                rec['dram_gbs'] = round(rec['dram_traffic_bytes'] / (rec['ms'] * 1e-3) / 1e9, 1)
        print(json.dumps(rec))
if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
