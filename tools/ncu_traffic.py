"""ncu report of one engine step (tools/profile_step.py) -> profiles/r02_traffic.json: per bench label, launches, average
duration and average DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch.  bench.py reads that file for
`roofline.traffic` (the numbers are measurements of a committed capture, not literals in bench.py).

    python tools/ncu_traffic.py gpurun_out/r02_step.ncu-rep profiles/r02_traffic.json [profiles/r02_step_kernels.txt]
"""
import csv, json, subprocess, sys

LABELS = [('gemm_tf32x3', 'gemm'), ('ego_encode_kernel', 'encode'), ('ego_rd', 'encode_rd'), ('bag_embed_fwd', 'bag_embed_fwd'),
          ('bag_embed_bwd_indexed', 'bag_embed_bwd_indexed'), ('bag_reduce', 'bag_embed_bwd_indexed'), ('bag_count', 'bag_index_build'),
          ('bag_fill', 'bag_index_build'), ('bag_scan', 'bag_index_build'), ('gine_fwd', 'gine_aggregate_fwd_ld'),
          ('gine_bwd', 'gine_aggregate_bwd_ld_noeps'), ('segment_pool_fwd', 'segment_pool_fwd'), ('segment_pool_bwd', 'segment_pool_bwd'),
          ('bn_act_fwd', 'bn_act_fwd'), ('bn_act_bwd', 'bn_act_bwd'), ('head_bn_linear_l1', 'head_bn_linear_l1'), ('adam', 'adam_step_device'),
          ('embedding_fwd', 'embedding_fwd'), ('embedding_bwd', 'embedding_bwd'), ('csr_', 'csr_build'), ('colsum', 'colsum')]

def unit_scale(u):
    return {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1)

txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, units = rows[0], rows[1]
col = {k: h.index(k) for k in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum')}
agg, lines = {}, []
for r in rows[2:]:
    name = r[col['Kernel Name']]
    us = float(r[col['gpu__time_duration.sum']]) * unit_scale(units[col['gpu__time_duration.sum']])
    rd = float(r[col['dram__bytes_read.sum']]) * unit_scale(units[col['dram__bytes_read.sum']])
    wr = float(r[col['dram__bytes_write.sum']]) * unit_scale(units[col['dram__bytes_write.sum']])
    label = next((lab for key, lab in LABELS if key in name), None)
    lines.append('%9.2f us  read %12.0f B  write %12.0f B  %-28s %s' % (us, rd, wr, label or '-', name[:110]))
    if label:
        a = agg.setdefault(label, dict(launches=0, us=0.0, rd=0.0, wr=0.0))
        a['launches'] += 1; a['us'] += us; a['rd'] += rd; a['wr'] += wr
out = {}
total_us = sum(a['us'] for a in agg.values())
for lab, a in agg.items():
    n = a['launches']
    out[lab] = dict(launches=n, avg_us=a['us'] / n, dram_read_bytes_per_launch=a['rd'] / n, dram_write_bytes_per_launch=a['wr'] / n,
                    dram_bytes_per_launch=(a['rd'] + a['wr']) / n, share_of_captured_kernel_time=a['us'] / total_us,
                    source='%s (ncu --set full --clock-control none, one eager step of config 2 / batch 256 on one stream, cold L2; '
                           'tools/profile_step.py + tools/ncu_traffic.py)' % 'profiles/r02_step_kernels.txt')
json.dump(out, open(sys.argv[2], 'w'), indent=1)
if len(sys.argv) > 3:
    open(sys.argv[3], 'w').write('\n'.join(lines) + '\n')
for lab, a in sorted(out.items(), key=lambda kv: -kv[1]['share_of_captured_kernel_time']):
    print('%-30s x%3d  %8.1f us avg  %12.0f dram B/launch  share %.3f' % (lab, a['launches'], a['avg_us'], a['dram_bytes_per_launch'], a['share_of_captured_kernel_time']))
