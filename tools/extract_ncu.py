"""Compact per-kernel summaries out of an `ncu --set full` report, for profiles/.

    python tools/extract_ncu.py gpurun_out/step_r02.ncu-rep profiles/step_ncu_full_r02.csv [kernel-name-substring]

Reads `ncu -i <rep> --page raw --csv` (ncu is installed in the build container; no GPU needed) and keeps the metrics
the roofline discussion in DESIGN.md cites, one row per (kernel, metric)."""
import csv
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
    'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
    'launch__shared_mem_per_block_static', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'metric', 'unit', 'value'])
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            name = d['Kernel Name']
            if want and want not in name:
                continue
            short = name.split('(')[0].replace('void ', '').replace('vb::', '').replace('fast::', '')
            for k in KEEP:
                if k in d and d[k] != '':
                    w.writerow([short, k, u[k], d[k]])


if __name__ == '__main__':
    main()
