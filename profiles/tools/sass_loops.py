"""Static picture of the loops of k_rx<8,1> (the span kernel of the receiver, QPSK slicer): instruction count and
opcode mix of every loop body of 100-700 SASS instructions.  Usage: python profiles/tools/sass_loops.py a.o [b.o ...]
(objects built with the flags of leansdr_b200/csrc/Makefile)."""
import re,collections,sys,subprocess
def analyze(obj):
    sass=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
    parts=sass.split('Function : ')
    for part in parts[1:]:
        name=part.split('\n',1)[0]
        if 'k_rxILi8ELi1E' not in name: continue
        ins=[]
        for l in part.splitlines():
            m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
            if m: ins.append((int(m.group(1),16),m.group(2).strip()))
        print(obj,len(ins),'instructions')
        for a,t in ins:
            m=re.search(r'BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)',t)
            if m:
                tgt=int(m.group(1),16)
                if tgt<a:
                    body=[u for x,u in ins if tgt<=x<=a]
                    if 100<=len(body)<700:
                        c=collections.Counter(re.sub(r'^@!?U?P\d+\s+','',u).split()[0].split('.')[0] for u in body)
                        print('  loop',hex(tgt),len(body),'LDG',c['LDG'],'STG',c['STG'],'BRA',c['BRA'],'IMAD',c['IMAD'],'MOV',c['MOV'],'FMUL',c['FMUL'])
for o in sys.argv[1:]: analyze(o)
