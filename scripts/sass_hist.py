"""SASS opcode histogram per kernel of a cuobjdump -sass listing (developer tool): sass_hist.py <listing> [name-substring]"""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
want = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if want not in name:
        continue
    ops = collections.Counter(m.group(1).split('.')[0] for m in
                              re.finditer(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', f))
    print(name[:80], sum(ops.values()))
    print('  ' + ', '.join(f"{k}:{v}" for k, v in ops.most_common(28)))
