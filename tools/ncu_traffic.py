"""Summarise an `ncu --set full` capture into profiles/r02_ncu_traffic.json (read by bench.py for roofline.traffic).

    python tools/ncu_traffic.py <capture.ncu-rep> <workload>:<n_gpus> [--out profiles/r02_ncu_traffic.json]

For every kernel in the capture: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured
launches), the launch duration under ncu and the capture file it came from."""
import csv
import io
import json
import os
import re
import subprocess
import sys


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = 'profiles/r02_ncu_traffic.json'
    if '--out' in sys.argv:
        out = sys.argv[sys.argv.index('--out') + 1]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    units = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    need = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum']
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'usecond': 1e3, 'nsecond': 1.0, 'msecond': 1e6}
    agg = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r'\(.*', '', r[col['Kernel Name']]).strip()
        name = name.split('::')[-1].split('<')[0]
        vals = []
        for m in need:
            v = float(r[col[m]].replace(',', ''))
            vals.append(v * scale.get(units[col[m]], 1.0))
        a = agg.setdefault(name, {'n': 0, 'rd': 0.0, 'wr': 0.0, 'ns': 0.0})
        a['n'] += 1; a['rd'] += vals[0]; a['wr'] += vals[1]; a['ns'] += vals[2]
    tab = {}
    if os.path.exists(out):
        tab = json.load(open(out))
    ent = tab.setdefault(key, {})
    for name, a in agg.items():
        ent[name] = {'dram_bytes': (a['rd'] + a['wr']) / a['n'], 'dram_read': a['rd'] / a['n'], 'dram_write': a['wr'] / a['n'],
                     'ncu_duration_us': a['ns'] / a['n'] / 1e3, 'launches_captured': a['n'],
                     'source': f'ncu --set full --clock-control none, {os.path.basename(rep)}'}
    json.dump(tab, open(out, 'w'), indent=1, sort_keys=True)
    for name, e in ent.items():
        print(f"{key} {name}: {e['dram_bytes'] / 1e6:.1f} MB/launch, {e['ncu_duration_us']:.1f} us")


if __name__ == '__main__':
    main()
