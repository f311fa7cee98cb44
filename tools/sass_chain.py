"""Estimate the issue time of a SASS loop for a lone warp: list scheduling is already done by ptxas, so walk the loop body
in order, charge 1 issue cycle per instruction (2.2 for fp64), and stall when a source register was written fewer than
`lat` cycles ago (fp64 8 cycles, others 4, LDS 29).  usage: python tools/sass_chain.py file.sass <kernel-substring>"""
import re
import sys

txt = open(sys.argv[1]).read()
funcs = re.split(r'\n\s*Function : ', txt)
for f in funcs[1:]:
    name = f.split('\n')[0]
    if sys.argv[2] not in name:
        continue
    ins = []
    for line in f.split('\n'):
        mm = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', line)
        if mm:
            ins.append((int(mm.group(1), 16), mm.group(2)))
    loops = []
    for a, full in ins:
        t = full.split()
        op = t[1] if t[0].startswith('@') else t[0]
        if op.startswith('BRA'):
            mt = re.search(r'0x([0-9a-f]+)', full)
            if mt and int(mt.group(1), 16) < a:
                loops.append((int(mt.group(1), 16), a))
    print(name[-60:])
    for s, e in loops:
        body = [(a, full) for a, full in ins if s <= a <= e]
        if len(body) < 40:
            continue
        ready = {}
        t = 0.0
        stall_tot = 0.0
        nfp = 0
        for a, full in body:
            toks = full.replace(',', ' ').split()
            if toks[0].startswith('@'):
                toks = toks[1:]
            op = toks[0]
            regs = [re.sub(r'[^R0-9]', '', x.split('.')[0]) for x in toks[1:] if re.match(r'^-?\|?R\d+', x)]
            fp64 = op.startswith(('DFMA', 'DMUL', 'DADD'))
            dst = regs[:1]
            src = regs[1:]
            if op.startswith(('STS', 'STG', 'ST.')):
                dst, src = [], regs
            need = t
            for r in src:
                n = int(r[1:])
                for rr in (n, n + 1) if fp64 else (n,):
                    need = max(need, ready.get(rr, 0.0))
            stall_tot += need - t
            t = need
            lat = 8.0 if fp64 else (29.0 if op.startswith('LDS') else (26.0 if op.startswith('SHFL') else 4.0))
            for r in dst:
                n = int(r[1:])
                width = 2 if (fp64 or '.64' in op) else (4 if '.128' in op else 1)
                for rr in range(n, n + width):
                    ready[rr] = t + lat
            t += 2.2 if fp64 else 1.0
            nfp += fp64
        print(f'  loop {hex(s)}-{hex(e)}: {len(body)} instr, {nfp} fp64, est. {t:.0f} cycles, of which dependency stalls {stall_tot:.0f}')
