"""Summarise an .ncu-rep (read on the CPU box): headline metrics per kernel + hot SASS regions.
usage: python profiles/ncu_summary.py <file.ncu-rep> [launch_index ...]"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    for r in rows[2:]:
        print('==', r[h.index('Kernel Name')][:100])
        for w in WANT:
            if w in h:
                print('   %-82s %s %s' % (w, r[h.index(w)], rows[1][h.index(w)]))


def hot(rep, idx, thresh=0.004):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(idx), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print('## hot SASS regions of', rows[0][1][:90])
    h = rows[1]
    ie, src, smp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    print('total warp instructions', tot)
    groups = []
    for k, r in enumerate(data):
        c, s = int(r[ie]), int(r[smp])
        if groups and groups[-1][0] == c:
            groups[-1][1] += 1; groups[-1][3] += s
        else:
            groups.append([c, 1, k, s])
    for c, nn, k, s in groups:
        if c * nn > tot * thresh:
            print('idx %4d n=%3d exec=%10d  share=%5.1f%% samples=%6d : %s' % (k, nn, c, 100 * c * nn / tot, s, data[k][src].strip()[:70]))
    return data


if __name__ == '__main__':
    rep = sys.argv[1]
    raw(rep)
    for i in sys.argv[2:]:
        hot(rep, int(i))
