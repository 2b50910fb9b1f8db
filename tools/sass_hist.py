"""Dev helper: opcode histogram of one kernel in libnfc_b200.so.  usage: python tools_sass.py <substr> [<substr2>...]"""
import re, subprocess, sys
from collections import Counter
so = "neuralfreeconvection.jl_b200/libnfc_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
parts = re.split(r'\n\s*Function : ', txt)
for p in parts[1:]:
    name = p.split('\n', 1)[0]
    if all(k in name for k in sys.argv[1:]):
        lines = [l for l in p.split('\n') if re.search(r'/\*[0-9a-f]{4,6}\*/', l)]
        ops = [re.search(r'\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l) for l in lines]
        c = Counter(m.group(1).split('.')[0] for m in ops if m)
        print(name, len(lines), "instructions")
        print(c.most_common(18))
        open('/tmp/kernel.sass', 'w').write(p)
