"""Compact per-launch summary of an .ncu-rep (read here with `ncu -i ... --page raw --csv`):
python tools/ncu_summary.py gpurun_out/prof_gemm_r1g.ncu-rep profiles/r01_ncu_full_conv_gemm_mb222.csv [names]"""
import csv
import io
import json
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('sm__cycles_elapsed.max', 'sm_cycles'), ('smsp__inst_executed.sum', 'warp_insts'),
        ('lts__t_bytes.sum', 'l2_bytes'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct')]
GEMM_NAMES = ['bank', 'conv2 IN', 'conv3', 'conv4 s2 IN+avg', 'conv5', 'conv6 s2 IN+avg', 'conv7', 'conv8 s2 IN+avg', 'dense1',
              'dense2 IN+res', 'dense3', 'dense4 IN+res', 'gx', 'linear', 'd.conv1 PS', 'd.conv2 IN+up2', 'd.conv3 PS',
              'd.conv4 IN+up2', 'd.conv5 PS', 'd.conv6 IN+up2', 'd.dense1', 'd.dense2 IN+res', 'd.dense3', 'd.dense4 IN+res',
              'd.gx', 'd.dense5', 'd.linear']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    names = GEMM_NAMES if len(sys.argv) > 3 and sys.argv[3] == 'gemm' else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    col = {h: i for i, h in enumerate(H)}
    kn = col['Kernel Name']
    scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'second': 1e6}
    recs = []
    for n, r in enumerate(rows[2:]):
        rec = {'launch': n, 'kernel': r[kn].split('(')[0].replace('void ', '')}
        if names and n < len(names):
            rec['layer'] = names[n]
        for m, short in WANT:
            if m in col:
                v = float(r[col[m]].replace(',', '')) if r[col[m]] not in ('', 'n/a') else None
                u = U[col[m]]
                if v is not None and u in scale and short in ('duration', 'dram_read', 'dram_write', 'l2_bytes'):
                    v *= scale[u]
                rec[short + ('_us' if short == 'duration' else '_bytes' if short in ('dram_read', 'dram_write', 'l2_bytes') else '')] = v
        recs.append(rec)
    keys = []
    for r in recs:
        for k in r:
            if k not in keys:
                keys.append(k)
    with open(out, 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(recs)
    tot_r = sum(r.get('dram_read_bytes') or 0 for r in recs)
    tot_w = sum(r.get('dram_write_bytes') or 0 for r in recs)
    tot_t = sum(r.get('duration_us') or 0 for r in recs)
    print(json.dumps({'launches': len(recs), 'dram_read_bytes': tot_r, 'dram_write_bytes': tot_w, 'duration_us_sum': tot_t}))


if __name__ == '__main__':
    main()
