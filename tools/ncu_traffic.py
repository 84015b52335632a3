"""Per-kernel DRAM traffic of an .ncu-rep (`ncu --set full` capture): launches, summed and per-launch
dram__bytes_read.sum + dram__bytes_write.sum and duration.  Output: one JSON object (bench.py reads the committed copy
under profiles/ for its `roofline.traffic` field)."""
import csv
import json
import re
import subprocess
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
TUNIT = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}


def main(paths):
    out = {}
    for path in paths:
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        kn, rd, wr, tm = (hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                                                 'gpu__time_duration.sum'))
        for r in rows[2:]:
            name = re.sub(r'\(.*', '', r[kn])
            name = re.sub(r'<unnamed>::|void ', '', name)
            name = re.sub(r'<.*', '', name)
            d = out.setdefault(name, dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, time_ms=0.0, source=[]))
            d['launches'] += 1
            d['dram_read_bytes'] += float(r[rd].replace(',', '')) * UNIT[units[rd]]
            d['dram_write_bytes'] += float(r[wr].replace(',', '')) * UNIT[units[wr]]
            d['time_ms'] += float(r[tm].replace(',', '')) * TUNIT[units[tm]]
            if path not in d['source']:
                d['source'].append(path)
    for d in out.values():
        d['traffic_bytes_per_launch'] = (d['dram_read_bytes'] + d['dram_write_bytes']) / d['launches']
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1:])
