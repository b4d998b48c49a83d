"""Print the innermost MUFU-bearing loops of a SASS dump as a pipe-usage string (S=MUFU.SQRT/RSQ, E=MUFU.EX2,
f=packed FP32, s=scalar FP32, l=LDS, L=local memory, .=other).   cuobjdump -sass -fun <mangled> lib.so | python tools/sass_loops.py"""
import re, sys
lines = [l for l in sys.stdin if re.search(r'/\*[0-9a-f]{4,5}\*/', l)]
ins = [re.sub(r'/\*[0-9a-f]+\*/', '', l).strip().split(';')[0] for l in lines]
addr = [int(re.search(r'/\*([0-9a-f]{4,5})\*/', l).group(1), 16) for l in lines]
for i, t in enumerate(ins):
    m2 = re.search(r'BRA.* 0x([0-9a-f]+)', t)
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt < addr[i] and tgt in addr:
            j = addr.index(tgt); body = ins[j:i + 1]
            mufu = sum('MUFU' in b for b in body)
            if 8 <= mufu <= 100 and len(body) < 600:
                st = ''
                for t2 in body:
                    op = t2.split()[0] if not t2.startswith('@') else t2.split()[1]
                    st += ('S' if ('SQRT' in t2 or 'RSQ' in t2) else 'E' if 'EX2' in t2 else 'f' if op in ('FFMA2', 'FMUL2', 'FADD2') else
                           's' if op in ('FFMA', 'FMUL', 'FADD') else 'l' if 'LDS' in t2 else 'L' if ('LDL' in t2 or 'STL' in t2) else '.')
                print(j, i, 'len', len(body), 'mufu', mufu, 'f2', st.count('f'), 'fs', st.count('s'), '\n   ', st)
