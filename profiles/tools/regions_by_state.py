"""regions_by_state.py -- join `nvdisasm -gi` (inline-aware line info) with an `ncu --page source --csv --print-source sass` dump and\naggregate executed instructions, stall samples and touched 128-byte instruction lines by PART of the scene-graph machine\n(line ranges of glome_gen.cuh).  usage: regions_by_state.py all_gi.dis src.csv <mangled-kernel-substring>"""
import re,csv,sys,collections
dis=sys.argv[1]; src=sys.argv[2]; kname=sys.argv[3]
lines=open(dis,errors='replace').read().splitlines()
start=None
for i,l in enumerate(lines):
    if l.startswith('//---') and '.text.' in l and kname in l: start=i;break
chain=[]; insn=[]
pend=[]
for l in lines[start+1:]:
    if l.startswith('//---') and '.text.' in l: break
    m=re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m:
        pend.append((m.group(1).split('/')[-1],int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0)))
        continue
    if re.search(r'/\*[0-9a-f]{4,}\*/\s+\S',l):
        if pend: chain=pend; pend=[]
        insn.append(chain)
def region(ch):
    # ch: list of (file,line,infile,inline) innermost first. collect all (file,line) frames
    frames=[]
    for f,l,f2,l2 in ch:
        frames.append((f,l))
        if f2: frames.append((f2,l2))
    # outermost frame in glome_gen.cuh
    gen=[(f,l) for f,l in frames if f=='glome_gen.cuh']
    cu=[(f,l) for f,l in frames if f=='glome_cuda.cu']
    if gen:
        l=gen[-1][1]
        for lo,hi,name in R:
            if lo<=l<=hi: return name
        return 'gen:%d'%l
    if cu: return 'kernel'
    return 'other:%s'%(frames[-1][0] if frames else '?')
R=[(101,201,'gq_inside'),(203,211,'inside_all'),(225,314,'gq_metainfo'),(423,455,'test_simple(noinline)'),(458,474,'fill_hit?'),(477,492,'inside_fast?'),
   (525,545,'qvm_start'),(549,572,'qvm_step prolog'),(573,603,'BRANCH'),(604,644,'LIST'),(645,671,'ENTER head+group'),(672,689,'ENTER bih'),(690,711,'ENTER mesh'),(712,739,'ENTER instance'),(740,766,'ENTER csg'),(767,802,'ENTER bound'),
   (803,844,'RET small'),(845,870,'RET inst'),(871,906,'RET misc'),(907,1012,'RET DIFF'),(1013,1145,'RET ISECT'),(1146,1162,'abort'),(1165,1170,'gq_query'),(1179,1288,'debug'),
   (1336,1343,'shm_start'),(1346,1624,'shm_step')]
rows=list(csv.reader(open(src)))
hdr=None;data=[]
for r in rows:
    if r and r[0]=='Address': hdr=r; continue
    if hdr and len(r)>=len(hdr)-2: data.append(r)
col={n:i for i,n in enumerate(hdr)}
def num(x):
    try: return float(x)
    except: return 0.0
n=min(len(data),len(insn))
print('sass',len(data),'dis',len(insn))
agg=collections.defaultdict(lambda:[0,0,0,0,0,0,set()])
T=[0,0,0,0]
for k in range(n):
    r=data[k]
    ie=num(r[col['Instructions Executed']]); te=num(r[col['Thread Instructions Executed']]); ss=num(r[col['# Samples']]); ni=num(r[col['stall_no_inst']]); lsb=num(r[col['stall_long_sb']])
    a=agg[region(insn[k])]
    a[0]+=ie;a[1]+=te;a[2]+=ss;a[3]+=ni;a[4]+=lsb;a[5]+=1
    if ie>0: a[6].add(int(r[col['Address']],16)//128 if r[col['Address']].startswith('0x') else int(r[col['Address']])//128)
    T[0]+=ie;T[1]+=te;T[2]+=ss;T[3]+=ni
print('%-24s %7s %7s %6s %8s %8s %8s %7s'%('region','winstr%','active','sass','samples%','noinst%','longsb%','linesKB'))
for k,a in sorted(agg.items(),key=lambda kv:-kv[1][2]):
    print('%-24s %6.2f%% %7.2f %6d %7.2f%% %7.2f%% %7.2f%% %7.1f'%(k,100*a[0]/T[0],a[1]/max(a[0],1),a[5],100*a[2]/T[2],100*a[3]/T[2],100*a[4]/T[2],len(a[6])*128/1024))
print('total winstr %.4g active %.2f'%(T[0],T[1]/T[0]))
