"""Per-source-line summary of an ncu report captured with --import-source on (-lineinfo build).
    python tools/ncu_source.py REPORT.ncu-rep [kernel-launch-index] [top]
Sums, per (file, line), the SASS rows of `ncu --page source --print-source cuda,sass --csv`: warp instructions executed,
thread instructions, stall samples; prints the top lines by warp instructions and by stall samples."""
import collections
import csv
import io
import subprocess
import sys


def load(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    kernels = []          # list of dict (file,line) -> [inst, thread_inst, samples, src]
    cur_file = None
    hdr = None
    agg = None
    last_fn = None
    key = None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == 'File Path':
            cur_file = row[1].split('/')[-1]
            continue
        if row[0] == 'Function Name':
            if row[1] != last_fn or agg is None:
                pass
            last_fn = row[1]
            continue
        if row[0] == 'Line No':
            hdr = row
            ii, ti, si = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
            if agg is None or (cur_file, 'start') in agg:
                agg = collections.OrderedDict()
                kernels.append((last_fn, agg))
            agg[(cur_file, 'start')] = None
            continue
        if hdr is None:
            continue
        if row[0] != '':
            key = (cur_file, int(row[0]))
            agg.setdefault(key, [0, 0, 0, row[1].strip()])
        elif key is not None and row[2] not in ('...', ''):
            try:
                a = agg[key]
                a[0] += int(row[ii]); a[1] += int(row[ti]); a[2] += int(row[si])
            except (ValueError, IndexError):
                pass
    return kernels


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    kernels = load(rep)
    print(len(kernels), 'kernel views')
    fn, agg = kernels[which]
    rows = [(k, v) for k, v in agg.items() if v]
    tot_i = sum(v[0] for _, v in rows); tot_t = sum(v[1] for _, v in rows); tot_s = sum(v[2] for _, v in rows)
    print(fn[:100]); print(f'warp inst {tot_i:,}  thread inst {tot_t:,}  avg lanes {tot_t / max(tot_i, 1):.2f}  samples {tot_s:,}')
    print('--- by warp instructions')
    for (f, l), v in sorted(rows, key=lambda kv: -kv[1][0])[:top]:
        print(f'{f:22s}:{l:4d} inst {v[0] / tot_i * 100:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  stall {v[2] / max(tot_s, 1) * 100:5.1f}%  | {v[3][:90]}')
    print('--- by stall samples')
    for (f, l), v in sorted(rows, key=lambda kv: -kv[1][2])[:top // 2]:
        print(f'{f:22s}:{l:4d} inst {v[0] / tot_i * 100:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  stall {v[2] / max(tot_s, 1) * 100:5.1f}%  | {v[3][:90]}')


if __name__ == '__main__':
    main()
