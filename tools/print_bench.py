import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print(sys.argv[1], 'value %.3gM ms %.3f gru %.3f gemm %.3f other %.3f e2e %.3gM'%(d['value']/1e6,d['ms_per_step'],r['gru_ms_per_step'],r['kernel_ms_per_step'],r['other_ms_per_step'],d['e2e']['value']/1e6), d.get('parity'))
