"""profiles/r02_sass_gemm.txt: per GEMM instantiation of libescgnn_b200.so the counts of the SASS mnemonics that prove the tcgen05 /
TMA / TMEM path (cuobjdump -sass, runs in the build container), plus an excerpt of the default forward kernel.
    python tools/sass_evidence.py > profiles/r02_sass_gemm.txt"""
import os, re, subprocess, sys
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(HERE, 'esc_gnn_b200', 'libescgnn_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
MN = ['UTCHMMA', 'UTMALDG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS']
print('SASS evidence for the tcgen05 / TMA / TMEM path of esc_gnn_b200/libescgnn_b200.so (cuobjdump -sass, sm_100a; this code state).')
print('Counts per kernel instantiation of the mnemonics B200_PROFILING.md names: UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load),')
print('LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  Template arguments of')
print('gemm_tf32x3_ts_kernel: <BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS, EPI, SW (splitter / epilogue warps), KBG (k-block groups)>.\n')
print('%-150s' % 'kernel' + ''.join('%9s' % m for m in MN))
cur, body, rows = None, [], []
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        if cur:
            rows.append((cur, body))
        cur, body = m.group(1), []
    elif cur:
        body.append(line)
if cur:
    rows.append((cur, body))
tot = dict.fromkeys(MN, 0)
default = None
for name, body in rows:
    if 'gemm_tf32x3' not in name:
        continue
    joined = '\n'.join(body)
    c = {m: len(re.findall(r'\b' + m, joined)) for m in MN}
    if not any(c.values()):
        continue
    for m in MN:
        tot[m] += c[m]
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'^void \(anonymous namespace\)::', '', dem)
    dem = re.sub(r'\(CUtensorMap_st.*$', '', dem)
    print('%-150s' % dem[:150] + ''.join('%9d' % c[m] for m in MN))
    if 'ts_kernel<128, false, false, 4, 4, 0, 8, 2>' in dem:
        default = (dem, body)
print('%-150s' % 'total' + ''.join('%9d' % tot[m] for m in MN))
if default:
    print('\nExcerpt of %s (node-level forward products of the training step): tensor-core, TMA and tensor-memory instructions in order\n' % default[0])
    for l in default[1]:
        if re.search(r'UTCHMMA|UTMALDG|LDTM|STTM|UTCBAR|UTMAPF|UTCATOM|ACQBULK|R2UR|UTCALLOC|TMEM', l):
            print('   ' + l.strip()[:170])
