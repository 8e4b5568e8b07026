"""Summarise an ncu report (raw page + source page) into a small text file for profiles/."""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_atom.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']


def main(rep, note=''):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print('# %s' % rep)
    if note:
        print('# ' + note)
    for r in rows[2:]:
        print('--- %s' % r[hdr.index('Kernel Name')][:70])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-78s %s %s' % (w, r[i][:24], units[i]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    agg, samp, text = collections.Counter(), collections.Counter(), {}
    cur, h = None, None
    for r in csv.reader(src.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            h = r
            continue
        if h is None or not r[0].isdigit():
            continue
        d = dict(zip(h, r))
        key = (cur, int(r[0]))
        def num(x):
            try:
                return int(x)
            except (TypeError, ValueError):
                return 0
        agg[key] += num(d.get('Instructions Executed'))
        samp[key] += num(d.get('# Samples'))
        text[key] = r[1][:100]
    ti, ts = sum(agg.values()) or 1, sum(samp.values()) or 1
    print('--- source hot spots over all captured launches (share of warp instructions / of stall samples)')
    for key, v in samp.most_common(16):
        print('  %-18s %4d inst=%5.1f%% samp=%5.1f%%  %s' % (key[0], key[1], 100 * agg[key] / ti, 100 * v / ts, text[key]))


if __name__ == '__main__':
    main(sys.argv[1], ' '.join(sys.argv[2:]))
