"""Print the judged metrics of `ncu --set full` reports (read here, on the CPU box) as a markdown table.

    python scripts/ncu_summary.py gpurun_out/prof_a.ncu-rep [gpurun_out/prof_b.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration under ncu'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput % of peak'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots active %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % of max'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]


def main(paths):
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print('\n### `%s`  (%s)\n' % (r[hdr.index('Kernel Name')][:100], p.split('/')[-1]))
            print('| metric | value |\n|---|---:|')
            for k, label in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print('| %s (`%s`) | %s %s |' % (label, k, r[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1:])
