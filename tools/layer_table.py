"""Per-layer table from an ncu launch list of `tools/profile_step.py B 2` (second repetition)."""
import csv, sys
path, B = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi = H.index('Kernel Name'), H.index('Metric Value')
data = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
fwd = [d for d in data if not any(s in d[0] for s in ('pack_weight', 'fold_bias', 'transpose_', 'gru_pack', 'reduce_kernel'))]
n = len(fwd) // 2
F = lambda co, ci, k, T: 2 * co * ci * k * T * B / 1e9
names = [('bank', 2 * 128 * 513 * 28 * 128 * B / 1e9), ('conv2 IN', F(512, 1409, 1, 128)), ('conv3', F(512, 512, 5, 128)), ('conv4 s2 IN+avg', F(512, 512, 5, 64)),
         ('conv5', F(512, 512, 5, 64)), ('conv6 s2 IN+avg', F(512, 512, 5, 32)), ('conv7', F(512, 512, 5, 32)), ('conv8 s2 IN+avg', F(512, 512, 5, 16)),
         ('dense1', F(512, 512, 1, 16)), ('dense2 IN+res', F(512, 512, 1, 16)), ('dense3', F(512, 512, 1, 16)), ('dense4 IN+res', F(512, 512, 1, 16)),
         ('gx', F(768, 512, 1, 16)), ('linear nct32', F(1024, 768, 1, 16)),
         ('d.conv1 PS', F(2048, 1024, 3, 16)), ('d.conv2 IN+up2', F(1024, 1024, 3, 32)), ('d.conv3 PS', F(2048, 1024, 3, 32)), ('d.conv4 IN+up2', F(1024, 1024, 3, 64)),
         ('d.conv5 PS', F(2048, 1024, 3, 64)), ('d.conv6 IN+up2', F(1024, 1024, 3, 128)), ('d.dense1', F(1024, 1024, 1, 128)), ('d.dense2 IN+res', F(1024, 1024, 1, 128)),
         ('d.dense3', F(1024, 1024, 1, 128)), ('d.dense4 IN+res', F(1024, 1024, 1, 128)), ('d.gx', F(3072, 1024, 1, 128)), ('d.dense5', F(1024, 2048, 1, 128)),
         ('d.linear nct32', F(513, 1024, 1, 128))]
gi = 0; tot = 0; ideal = 0; other = 0
for k, t in fwd[n:]:
    if 'conv_gemm' in k:
        nm, gf = names[gi]; gi += 1
        idu = gf / 1368 * 1e3
        print(f"{nm:18s} {t / 1000:8.1f} us  ideal@1368TF {idu:7.1f} us  eff {100 * idu / (t / 1000):5.1f}%")
        tot += t / 1000; ideal += idu
    else:
        print(f"   [{k[:44]}] {t / 1000:8.1f} us"); other += t / 1000
print(f'gemm total {tot:.1f} us, ideal {ideal:.1f} us, eff {100 * ideal / tot:.1f}% ; non-gemm {other:.1f} us')
