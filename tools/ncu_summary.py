"""Prints the handful of counters we judge a kernel by from an `ncu --set full` report (run where ncu is installed):
duration, DRAM bytes, DRAM / tensor-pipe / issue utilisation, occupancy limits and the top warp-stall reasons.
Usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',  # tcgen05 pipe busy (matches MMA count x tools/mma_probe cost)
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('kernel:', d.get('Kernel Name'))
        for i, k in enumerate(hdr):
            if k in WANT or k.endswith('tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'):
                print(f'  {k:88s} {d[k]} {units[i]}')
        st = [(float(d[k].replace(',', '')), k) for k in hdr
              if k.startswith('smsp__pcsamp_warps_issue_stalled') and not k.endswith('not_issued') and d[k] not in ('', 'n/a')]
        tot = sum(v for v, _ in st)
        print('  warp stall reasons (pc sampling):')
        for v, k in sorted(st, reverse=True)[:8]:
            print(f'    {k[33:]:40s} {100 * v / max(tot, 1):5.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
