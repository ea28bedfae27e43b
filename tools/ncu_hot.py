"""Print headline metrics and the hottest SASS lines (by stall samples) of every kernel in an .ncu-rep file."""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_op_read.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum']


def run(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, top=25):
    raw = run(rep, 'raw')
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in raw[2:]:
        print('====', r[idx['Kernel Name']][:100])
        for w in WANT:
            if w in idx:
                print(f"  {w:75s} {r[idx[w]]} {units[idx[w]]}")
        for h in hdr:
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                v = float(r[idx[h]] or 0)
                if v > 0.3:
                    print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
    src = run(rep, 'source')
    h = None
    body = []
    for r in src:
        if r and r[0] == 'Kernel Name':
            if body:
                break
            continue
        if r and r[0] == 'Address':
            h = r
            continue
        if h and len(r) == len(h):
            body.append(r)
    if not body:
        return
    si, ii = h.index('# Samples'), h.index('Instructions Executed')
    tot = sum(int(r[si]) for r in body)
    print(f"-- first kernel: {tot} samples over {len(body)} SASS lines; hottest:")
    for i, r in sorted(sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:top]):
        print(f"  {i:5d} {r[1].strip()[:70]:70s} samples={r[si]:>5s} exec={r[ii]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
