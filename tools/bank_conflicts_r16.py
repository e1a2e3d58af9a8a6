# bank conflicts for the radix-16 plan: M = 16*16*L, T = M/16 threads, 16 points per thread
import itertools
def wf64(addrs):
    tot=0
    for h in range(0,len(addrs),16):
        lanes=addrs[h:h+16]; banks={}
        for a in lanes: banks.setdefault(a%16,set()).add(a)
        tot+=max(len(s) for s in banks.values())
    return tot
def wf128(addrs):
    tot=0
    for h in range(0,len(addrs),8):
        lanes=addrs[h:h+8]; banks={}
        for a in lanes: banks.setdefault((a//2)%8,set()).add(a)
        tot+=max(len(s) for s in banks.values())
    return tot
def ev(M,pad):
    T=M//16; L=M//256
    nw=max(1,T//32); wl=min(T,32)
    res=[]
    # stage A
    w=n=0
    for warp in range(nw):
        for j in range(16):
            w+=wf64([pad(warp*32+l + T*j) for l in range(wl)]); n+=1
    res.append(('A',w/n*32/wl))
    w=n=0
    for warp in range(nw):
        for j in range(16):
            a=[]
            for l in range(wl):
                t=warp*32+l; blk=t//L; i=t%L
                a.append(pad(blk*16*L+i+L*j))
            w+=wf64(a); n+=1
    res.append(('B',w/n*32/wl))
    w=n=0
    for warp in range(nw):
        for q in range(8):
            w+=wf128([pad(16*(warp*32+l)+2*q) for l in range(wl)]); n+=1
    res.append(('C128',w/n*32/wl))
    return res
cands={'none':lambda e:e,'2(e>>4)':lambda e:e+2*(e>>4)}
for a,b,c,d in itertools.product([0,2],[0,2,4],[0,2,4,8],[0,2,4,8]):
    cands[f'2(e>>4)+{a}(e>>5)+{b}(e>>6)+{c}(e>>7)+{d}(e>>8)']=(lambda a,b,c,d:(lambda e:e+2*(e>>4)+a*(e>>5)+b*(e>>6)+c*(e>>7)+d*(e>>8)))(a,b,c,d)
best=[]
for name,p in cands.items():
    tot=0; rows=[]
    for M in (512,1024,2048,4096):
        r=ev(M,p); rows.append((M,r)); tot+=sum(x for _,x in r)
    best.append((tot,name,rows))
best.sort(key=lambda x:x[0])
for tot,name,rows in best[:6]+[b for b in best if b[1] in('none','2(e>>4)')]:
    print(round(tot,1),name)
    for M,r in rows: print('   ',M,[(k,round(v,2)) for k,v in r])
