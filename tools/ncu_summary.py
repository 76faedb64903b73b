"""Helpers to read .ncu-rep files brought back from the GPU box (no GPU needed here).

    python tools/ncu_summary.py summary gpurun_out/prof.ncu-rep          # key metrics per launch
    python tools/ncu_summary.py blocks  gpurun_out/prof.ncu-rep [n]      # hottest SASS blocks by executed instr
    python tools/ncu_summary.py sass    gpurun_out/prof.ncu-rep lo hi    # SASS lines [lo,hi) with counts + stalls
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summary(rep, out=None):
    rows = ncu_csv(rep, 'raw')
    hdr, units = rows[0], rows[1]
    lines = []
    ki = hdr.index('Kernel Name')
    lines.append(['metric', 'unit'] + [r[ki].split('(')[0][:50] for r in rows[2:]])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append([k, units[i]] + [r[i] for r in rows[2:]])
    if out:
        with open(out, 'w') as f:
            csv.writer(f).writerows(lines)
    for l in lines:
        print(f"{l[0]:82s} {l[1]:16s} " + ' '.join(f"{x:>16s}" for x in l[2:]))


def sass_rows(rep):
    rows = ncu_csv(rep, 'source')
    hdr = rows[1]
    si, ii = hdr.index('Source'), hdr.index('Instructions Executed')
    sti = hdr.index('Warp Stall Sampling (All Samples)')
    ins = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        if len(r) > ii:
            try:
                ins.append((r[si].strip(), int(r[ii]), int(r[sti] or 0)))
            except ValueError:
                pass
    return ins


def blocks(rep, n=25):
    ins = sass_rows(rep)
    tot = sum(x[1] for x in ins)
    tots = sum(x[2] for x in ins)
    print(len(ins), 'SASS instructions; executed (warp-level):', tot, '; stall samples:', tots)
    bl, cur = [], None
    for idx, (s, c, st) in enumerate(ins):
        if cur and abs(c - cur['n']) <= 0.02 * max(cur['n'], 1) + 1:
            cur['end'] = idx; cur['sum'] += c; cur['st'] += st
        else:
            if cur:
                bl.append(cur)
            cur = {'start': idx, 'end': idx, 'n': c, 'sum': c, 'st': st}
    bl.append(cur)
    bl.sort(key=lambda b: -b['sum'])
    for b in bl[:n]:
        ops = {}
        for s, c, st in ins[b['start']:b['end'] + 1]:
            op = s.split()[0] if not s.startswith('@') else s.split()[1]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
        print(f"[{b['start']:5d}-{b['end']:5d}] len={b['end']-b['start']+1:4d} x{b['n']:9d} = {b['sum']:11d} {100*b['sum']/tot:5.1f}% "
              f"stall {100*b['st']/max(tots,1):5.1f}%  {top}")


def sass(rep, lo, hi):
    ins = sass_rows(rep)
    for i in range(lo, min(hi, len(ins))):
        print(i, str(ins[i][1]).rjust(9), str(ins[i][2]).rjust(6), ins[i][0])


if __name__ == '__main__':
    cmd, rep = sys.argv[1], sys.argv[2]
    if cmd == 'summary':
        summary(rep, sys.argv[3] if len(sys.argv) > 3 else None)
    elif cmd == 'blocks':
        blocks(rep, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    elif cmd == 'sass':
        sass(rep, int(sys.argv[3]), int(sys.argv[4]))
