"""Regenerate profiles/traffic.json from the committed ncu raw pages (``ncu -i <rep> --page raw --csv``).

    python tools/make_traffic_json.py profiles/r02_fd_tma_512x512x512_raw.csv:512x512x512 \
                                      profiles/r02_fd_tma_128x2048x2048_raw.csv:128x2048x2048 ...

Every argument is ``<raw csv>:<per-GPU shape>``; bench.py looks its per-GPU shape up under ``shapes`` and copies
``dram_bytes_per_launch`` into ``roofline.traffic``.  Per launch = mean over the captured launches of
``dram__bytes_read.sum + dram__bytes_write.sum``; the units row of the raw page is honoured."""
import csv
import json
import os
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'usecond': 1.0,
        'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6}


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name):
        i = col[name]
        return [float(r[i].replace(',', '')) * UNIT[units[i]] for r in data]
    rd, wr, us = get('dram__bytes_read.sum'), get('dram__bytes_write.sum'), get('gpu__time_duration.sum')
    n = len(data)
    out = {'kernel': data[0][col['Kernel Name']], 'launches_captured': n, 'grid_size': data[0][col['Grid Size']] if 'Grid Size' in col else None,
           'dram_bytes_read_per_launch': sum(rd) / n, 'dram_bytes_write_per_launch': sum(wr) / n,
           'dram_bytes_per_launch': (sum(rd) + sum(wr)) / n, 'gpu_time_duration_us': sum(us) / n}
    if 'launch__registers_per_thread' in col:
        out['registers_per_thread'] = int(float(data[0][col['launch__registers_per_thread']]))
    return out


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shapes = {}
    for arg in sys.argv[1:]:
        path, shape = arg.rsplit(':', 1)
        d = read_raw(path)
        nodes = 1
        for v in shape.split('x'):
            nodes *= int(v)
        bpn = float(os.environ.get('PDX_BYTES_PER_NODE', '32'))
        d['algorithmic_bytes_per_launch'] = nodes * bpn
        d['traffic_over_algorithmic'] = d['dram_bytes_per_launch'] / d['algorithmic_bytes_per_launch']
        d['source'] = os.path.relpath(path, root) + ' (ncu --set full --clock-control none --import-source on, tools/ncu_target.py --shape %s)' % shape.replace('x', ',')
        shapes[shape] = d
    out = {'generated_by': 'tools/make_traffic_json.py ' + ' '.join(os.path.relpath(a.rsplit(':', 1)[0], root) + ':' + a.rsplit(':', 1)[1] for a in sys.argv[1:]),
           'shapes': shapes}
    with open(os.path.join(root, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (round(v['dram_bytes_per_launch'] / 1e9, 4), round(v['traffic_over_algorithmic'], 4)) for k, v in shapes.items()}))


if __name__ == '__main__':
    main()
