#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per launch: duration, DRAM bytes and
throughput, tensor-pipe %, achieved occupancy, registers.  Usage: python scripts/ncu_summary.py rep.ncu-rep [out.md]"""
import csv
import io
import re
import subprocess
import sys

COLS = {
    'dur_us': 'gpu__time_duration.sum',
    'dram_rd': 'dram__bytes_read.sum',
    'dram_wr': 'dram__bytes_write.sum',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'tensor_pct2': 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'occ': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'regs': 'launch__registers_per_thread',
    'l2_rd': 'lts__t_bytes.sum',
}


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    u = unit.lower()
    mult = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)
    return v * mult


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ['| # | kernel | grid | dur us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 % | tensor % | SM % | occ % | regs |',
             '|---|---|---|---|---|---|---|---|---|---|---|---|---|']
    for r in body:
        name = re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').replace('tok::', '')

        def g(key):
            c = COLS[key]
            return (r[idx[c]], units[idx[c]]) if c in idx else ('nan', '')
        dur, du = g('dur_us')
        dur = float(dur.replace(',', ''))
        dur_us = dur / 1e3 if du in ('ns', 'nsecond') else (dur if du in ('us', 'usecond') else dur * 1e3)
        rd = to_bytes(*g('dram_rd'))
        wr = to_bytes(*g('dram_wr'))
        tp = g('tensor_pct')[0]
        lines.append(f"| {r[idx['ID']]} | {name[:44]} | {r[idx['Grid Size']]} | {dur_us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                     f"{(rd + wr) / dur_us / 1e3:.0f} | {g('dram_pct')[0]} | {g('l2_pct')[0]} | {tp} | {g('sm_pct')[0]} | "
                     f"{g('occ')[0]} | {g('regs')[0]} |")
    out = '\n'.join(lines)
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(out + '\n')


if __name__ == '__main__':
    main()
