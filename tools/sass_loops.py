import re,sys
from collections import Counter
ins=[]
for l in open(sys.argv[1]):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2)))
addr={a:i for i,(a,_) in enumerate(ins)}
def op(t):
    p=t.split()
    return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
for i,(a,t) in enumerate(ins):
    if 'BRA' in t:
        m2=re.search(r'0x([0-9a-f]+)',t)
        if m2:
            tgt=int(m2.group(1),16)
            if tgt<a and tgt in addr:
                body=ins[addr[tgt]:i+1]
                c=Counter(op(x[1]) for x in body)
                print(hex(tgt),hex(a),len(body),dict(c.most_common(8)))
if len(sys.argv)>3:
    lo,hi=int(sys.argv[2],16),int(sys.argv[3],16)
    for a,t in ins:
        if lo<=a<hi: print(hex(a),t)
