#!/usr/bin/env python
"""cuobjdump -sass of the built library: one listing per hot kernel + a mnemonic histogram (which memory / async-copy /
shuffle instructions the kernels really contain).  usage: tools/sass_excerpts.py <lib.so> <output prefix>"""
import collections
import re
import subprocess
import sys

FULL = ['k_walk_pair<false>', 'k_walk_pair<(bool)0>', 'k_fix_chain', 'k_emit_bm_fast', 'k_lf_bm', 'k_lf_t1', 'k_lf_tma']  # full listings (the rest: summary only)
HOT = ['k_walk_pair', 'k_fix_chain', 'k_emit_bm_fast', 'k_write_walk', 'k_fine_walk', 'k_list_rank', 'k_scatter_ka', 'k_lf_bm', 'k_lf_t1', 'k_lf_tma',
       'k_fmd_payload', 'k_lf<']


def main(lib, prefix):
    out = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    archs = sorted(set(re.findall(r'arch = (sm_\w+)', out)))
    funcs = re.split(r'\n\s*Function : ', out)[1:]
    summary = ['cuobjdump -sass %s   (cubin architectures: %s)' % (lib, ', '.join(archs)), '']
    with open(prefix + '_hot_kernels.txt', 'w') as full:
        for f in funcs:
            name = f.split('\n', 1)[0].strip()
            dem = subprocess.run(['c++filt', name], stdout=subprocess.PIPE).stdout.decode().strip()
            if not any(h in dem for h in HOT):
                continue
            ins = re.findall(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', f)
            hist = collections.Counter(i.split('.')[0] for i in ins)
            mem = collections.Counter(i for i in ins if i.split('.')[0] in ('LDG', 'STG', 'LDS', 'STS', 'LDGSTS', 'UBLKCP', 'SYNCS', 'SHFL', 'POPC', 'ATOMG', 'RED', 'LDGDEPBAR', 'DEPBAR', 'BAR', 'VOTE', 'MATCH', 'LDL', 'STL'))
            summary.append('%s' % dem[:150])
            summary.append('   %d instructions; top: %s' % (len(ins), ', '.join('%s %d' % kv for kv in hist.most_common(8))))
            summary.append('   memory / sync: %s' % ', '.join('%s %d' % kv for kv in sorted(mem.items())))
            if any(h in dem for h in FULL):
                body = '\n'.join(re.sub(r'\s*/\* 0x[0-9a-f]{16} \*/', '', ln).rstrip() for ln in f.split('\n') if not re.match(r'^\s*/\* 0x[0-9a-f]{16} \*/\s*$', ln))
                full.write('==== %s\n%s\n' % (dem, body))
    open(prefix + '_summary.txt', 'w').write('\n'.join(summary) + '\n')
    print('\n'.join(summary[:40]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
