"""Extract the metrics the roofline cites from ncu --set full reports (raw page)."""
import csv
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'smsp__inst_executed_op_global_red.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def main(paths):
    traffic = {}
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f'== {p}\n   kernel: {d.get("Kernel Name")}   grid {d.get("Grid Size")} block {d.get("Block Size")}')
            for k in WANT:
                if k in d:
                    print(f'   {k:66s} {d[k]:>18s} {u[k]}')
            try:
                rd = float(d['dram__bytes_read.sum']) * UNIT[u['dram__bytes_read.sum']]
                wr = float(d['dram__bytes_write.sum']) * UNIT[u['dram__bytes_write.sum']]
                print(f'   => DRAM traffic per launch: {(rd + wr) / 1e9:.2f} GB')
                traffic[d.get('Kernel Name', p)] = rd + wr
            except Exception:
                pass
    return traffic


if __name__ == '__main__':
    t = main(sys.argv[1:])
    print(json.dumps(t))
