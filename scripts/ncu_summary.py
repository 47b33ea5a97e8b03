"""Summarise an .ncu-rep (raw page) into the handful of metrics we steer by."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fmaheavy.avg.pct', 'sm__pipe_fmaheavy_cycles_active.avg.pct',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__pipe_alu_cycles_active.avg.pct', 'smsp__issue_active.avg.pct', 'sm__throughput.avg.pct',
        'smsp__inst_executed.sum ', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp', 'smsp__average_warps_issue_stalled',
        'sm__inst_executed_pipe_fma.avg.pct', 'sm__cycles_elapsed.avg ', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'sm__inst_executed_pipe_lsu.avg.pct',
        'smsp__cycles_active.avg ', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fmalite', 'smsp__inst_executed_op_shared', 'sm__cycles_active.avg ',
        'gpc__cycles_elapsed.avg.per_second', 'sm__inst_executed_pipe_uniform', 'sm__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_cbu', 'sm__inst_executed_pipe_adu']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('---', vals[hdr.index('Kernel Name')][:80])
        for h, u, v in zip(hdr, units, vals):
            if any((w in h + ' ') for w in WANT):
                print(f'  {h} [{u}] = {v}')
if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
