"""Print the key metrics of an .ncu-rep (raw page) for each profiled launch."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max']


def main(path, extra=()):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('--- ', r[hdr.index('Kernel Name')][:80])
        for k in list(KEYS) + list(extra):
            if k in hdr:
                i = hdr.index(k)
                print('   %-70s %s %s' % (k, r[i], units[i]))
        for i, k in enumerate(hdr):
            if 'warp_issue_stalled' in k and k.endswith('per_warp_active.pct'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 4:
                    print('   stall %-62s %.1f' % (k.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
