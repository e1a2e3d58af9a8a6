# smem bank-conflict estimate for the in-place radix-8 FFT exchange patterns
import itertools, sys
def wavefronts64(addrs):  # addrs: 32 float2 indices; half-warp phases
    tot=0
    for h in range(2):
        lanes=addrs[16*h:16*h+16]
        banks={}
        for a in lanes:
            banks.setdefault(a%16,set()).add(a)
        tot+=max(len(s) for s in banks.values())
    return tot
def wavefronts128(addrs):  # addrs: 32 float4 starts in float2 units (even), quarter-warp phases
    tot=0
    for h in range(4):
        lanes=addrs[8*h:8*h+8]
        banks={}
        for a in lanes:
            banks.setdefault((a//2)%8,set()).add(a)
        tot+=max(len(s) for s in banks.values())
    return tot
def stages(M):
    S=M; out=[]
    while S>=8:
        out.append(S); S//=8
    return out,S
def evaluate(M,pad):
    T=M//8
    st,tail=stages(M)
    res=[]
    for S in st:
        s=S//8
        w=0;n=0
        for warp in range(T//32):
            for j in range(8):
                addrs=[]
                for lane in range(32):
                    t=warp*32+lane
                    blk=t//s; i=t%s
                    addrs.append(pad(blk*S+i+s*j))
                w+=wavefronts64(addrs); n+=1
        res.append((S,w/n))
    # contiguous-8 stage with 128-bit accesses
    w=0;n=0
    for warp in range(T//32):
        for q in range(4):
            addrs=[pad(8*(warp*32+lane)+2*q) for lane in range(32)]
            w+=wavefronts128(addrs); n+=1
    res.append(('c8x128',w/n))
    return res
for M in (512,1024,2048,4096,8192):
    print(M,'nopad',evaluate(M,lambda e:e))
    for name,pad in [('e+e>>5*2',lambda e:e+2*(e>>5)),('e+e>>4*2',lambda e:e+2*(e>>4)),('e+(e>>3)*2',lambda e:e+2*(e>>3)),
                     ('e+2(e>>3)+2(e>>6)',lambda e:e+2*(e>>3)+2*(e>>6)),('e+2(e>>3)+2(e>>7)',lambda e:e+2*(e>>3)+2*(e>>7))]:
        print(M,name,evaluate(M,pad))
