"""Turn the raw outputs of tools/gpurun/gpu_final_r4.sh (gpurun_out/<prefix>_*) into the committed evidence under profiles/:
launch list summary, per-launch DRAM traffic table + JSON (what bench.py's roofline.traffic quotes, with the commit hash),
`ncu --set full` summary of the Gram capture, pytest / smoke output.   usage: python tools/make_evidence.py f4 r02"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import launch_summary  # noqa: E402
import ncu_summary  # noqa: E402
import ncu_stalls  # noqa: E402

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
TUNIT = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}


def main(prefix, rnd):
    go = os.path.join(ROOT, 'gpurun_out')
    pr = os.path.join(ROOT, 'profiles')
    commit = open(os.path.join(go, prefix + '_commit.txt')).read().strip()
    # launch list
    buf = io.StringIO()
    with redirect_stdout(buf):
        launch_summary.main(os.path.join(go, prefix + '_launches.csv'))
    with open(os.path.join(pr, rnd + '_launches_final.txt'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 1 --warmup 1 --no-e2e '
                '--no-cpu-baseline` (two cfg4 fits), commit %s (tools/gpurun/gpu_final_r4.sh); times under ncu are serialised and '
                'cold-cache: shares, not absolutes\n' % commit)
        f.write(buf.getvalue())
    # traffic
    with open(os.path.join(go, prefix + '_traffic.csv')) as f:
        lines = [l for l in f if not l.startswith('==')]
    per = {}
    for row in csv.DictReader(lines):
        d = per.setdefault(int(row['ID']), dict(kernel=row['Kernel Name'], grid=row['Grid Size']))
        v = float(row['Metric Value'].replace(',', ''))
        if row['Metric Name'] == 'gpu__time_duration.sum':
            d['ms'] = v * TUNIT[row['Metric Unit']]
        elif row['Metric Name'] == 'dram__bytes_read.sum':
            d['rd'] = v * UNIT[row['Metric Unit']]
        elif row['Metric Name'] == 'dram__bytes_write.sum':
            d['wr'] = v * UNIT[row['Metric Unit']]
    out = {'_commit': commit,
           '_command': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none '
                       '-k regex:"basis_kernel|gram_kernel" python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline '
                       '(tools/gpurun/gpu_final_r4.sh): the first cfg4 fit of the process'}
    with open(os.path.join(pr, rnd + '_ncu_traffic_launches.txt'), 'w') as f:
        f.write('# per-launch DRAM traffic of every K1 / K2 launch of one cfg4 fit (commit %s); times under ncu are cold-cache '
                'and serialised\n' % commit)
        f.write('%-4s %-34s %-16s %10s %12s %10s\n' % ('id', 'kernel', 'grid', 'read GB', 'write GB', 'ms'))
        for i, k in enumerate(sorted(per)):
            d = per[k]
            name = re.sub(r'\(.*', '', d['kernel'])
            name = re.sub(r'<unnamed>::|void ', '', name)
            f.write('%-4d %-34s %-16s %10.3f %12.3f %10.3f\n' % (i, name, d['grid'], d['rd'] / 1e9, d['wr'] / 1e9, d['ms']))
            base = re.sub(r'<.*', '', name)
            a = out.setdefault(base, dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, time_ms=0.0))
            a['launches'] += 1
            a['dram_read_bytes'] += d['rd']
            a['dram_write_bytes'] += d['wr']
            a['time_ms'] += d['ms']
    for k, a in out.items():
        if isinstance(a, dict):
            a['traffic_bytes_per_launch'] = (a['dram_read_bytes'] + a['dram_write_bytes']) / a['launches']
    with open(os.path.join(pr, rnd + '_ncu_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    # ncu --set full of the Gram capture
    rep = os.path.join(go, prefix + '_prof_gram.ncu-rep')
    buf = io.StringIO()
    with redirect_stdout(buf):
        ncu_summary.main(rep)
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
        tmp = '/tmp/_src.csv'
        open(tmp, 'w').write(src)
        ncu_stalls.main(tmp)
    with open(os.path.join(pr, rnd + '_ncu_gram.txt'), 'w') as f:
        f.write('# ncu --set full --clock-control none of the 12th Gram launch of a cfg4 fit (C = 168 new columns, the widest '
                'model), commit %s (tools/gpurun/gpu_final_r4.sh)\n' % commit)
        f.write(buf.getvalue())
    with open(os.path.join(pr, rnd + '_pytest_gpu.txt'), 'w') as f:
        f.write('# python -c "import __graft_entry__ as g; g.smoke()" and python -m pytest tests -m gpu -q on a B200, commit %s\n' % commit)
        f.write(open(os.path.join(go, prefix + '_smoke.log')).read())
        f.write(open(os.path.join(go, prefix + '_pytest.log')).read())
    print('evidence written for commit', commit)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
