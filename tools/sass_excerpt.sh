#!/bin/bash
# SASS evidence for the hot kernels of the config-2 module (runs on the CPU box: cuobjdump only needs the .so):
# instruction counts per kernel and the lines that show the fp64 tensor-core MMA (DMMA), 128-bit shared / global
# accesses and async copies.
so=$(ls hilo_mpc_b200/_gen/nmpc_cstr_klatt_engell_collocation_N50_*.so | head -1)
out=profiles
cuobjdump -sass $so > /tmp/cstr_sass.txt
python - <<'PY'
import re, collections
txt = open('/tmp/cstr_sass.txt').read()
parts = re.split(r'\n\s*Function : ', txt)
for part in parts[1:]:
    name = part.split('\n', 1)[0].strip()
    short = None
    for key, tag in (('eval_quad_kernel', 'eval_quad'), ('riccati6_kernel', 'riccati6'), ('step_quad_kernel', 'step_quad')):
        if key in name and (tag != 'riccati6' or 'Li2ELi6' in name):
            short = tag
    if short is None:
        continue
    ins = re.findall(r'/\*[0-9a-f]{4,5}\*/\s+(.*?);', part)
    ops = collections.Counter(i.split()[0] if not i.startswith('@') else i.split()[1] for i in ins)
    with open(f'profiles/sass_r2_{short}.txt', 'w') as f:
        f.write(f'# {name}\n# {len(ins)} SASS instructions; opcode histogram (top 25):\n')
        for op, n in ops.most_common(25):
            f.write(f'#   {n:6d}  {op}\n')
        f.write('# excerpts (first occurrences):\n')
        for pat in ('DMMA', 'LDS.128', 'STS.128', 'LDG.E.128', 'LDG.E.64', 'LDGSTS', 'SHFL', 'MUFU.RCP64H', 'DFMA', 'CCTL', 'LDC'):
            hits = [i for i in ins if pat in i][:3]
            for h in hits:
                f.write(f'{pat:12s} {h}\n')
    print(short, len(ins), ops.most_common(6))
PY
