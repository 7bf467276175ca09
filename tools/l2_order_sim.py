#!/usr/bin/env python
"""LRU model of the L2 for the child-tile reads of one wave of the NNNNANNNN tile lattice under different claim orders (CPU only).
python tools/l2_order_sim.py WAVE CAPACITY_IN_TILES   (wave 7, 4000 tiles reproduces the measured 58 % miss rate of the lexicographic order;
interleaving the top digit's segments gives 59-64 %, class-blocked orders 61-69 %: nothing beats ascending tile numbers)"""
import sys, itertools
from collections import OrderedDict
MASK=[1,2,4,8,5,10,6,9,12,3,14,13,11,7,15]
pc=lambda m: bin(m).count('1')
LEV=[pc(m)-1 for m in MASK]
SPL={4:[(0,2)],5:[(1,3)],6:[(2,1)],7:[(0,3)],8:[(2,3)],9:[(0,1)],10:[(1,8),(2,5),(3,6)],11:[(0,8),(2,7),(3,4)],12:[(0,5),(1,7),(3,9)],13:[(0,6),(1,4),(2,9)],14:[(6,7),(8,9),(4,5),(0,10),(1,11),(2,12),(3,13)]}
for d in range(4): SPL[d]=[]
W=[15**i for i in range(5)]
def tiles_of_wave(w):
    out=[]
    for ds in itertools.product(range(15),repeat=5):   # ds[0]=d1 (fastest) ... ds[4]=d5
        if sum(LEV[d] for d in ds)==w: out.append(ds)
    return out
def tid(ds): return sum(d*W[i] for i,d in enumerate(ds))
def children(ds):
    res=[]
    for i,d in enumerate(ds):
        for c1,c2 in SPL[d]:
            a=list(ds); a[i]=c1; res.append(tid(a))
            b=list(ds); b[i]=c2; res.append(tid(b))
    return res
def simulate(order, cap):
    cache=OrderedDict(); miss=0; tot=0
    for ds in order:
        for c in children(ds):
            tot+=1
            if c in cache: cache.move_to_end(c)
            else:
                miss+=1; cache[c]=1
                if len(cache)>cap: cache.popitem(last=False)
    return miss,tot
w=int(sys.argv[1]); cap=int(sys.argv[2])
T=tiles_of_wave(w)
lex=sorted(T,key=tid)
print("wave",w,"tiles",len(T))
m,t=simulate(lex,cap); print("lex",m/t)
# d5-interleaved with block B: within each level class of d5, interleave segments
def interleave(B):
    segs={}
    for ds in lex: segs.setdefault(ds[4],[]).append(ds)
    out=[]
    bylev={}
    for d5,s in segs.items(): bylev.setdefault(LEV[d5],[]).append(d5)
    for lv in sorted(bylev):
        ds5=sorted(bylev[lv]); n=len(segs[ds5[0]])
        for st in range(0,n,B):
            for d5 in ds5: out.extend(segs[d5][st:st+B])
    return out
for B in (1,8,32,128,512,2048):
    o=interleave(B); assert len(o)==len(T)
    m,t=simulate(o,cap); print("interleave d5 B=%d"%B, m/t)
# interleave both d5 and d4 classes: sort key (lev5,lev4, rest3..., d4, d5)
def key2(ds): return (LEV[ds[4]],LEV[ds[3]], ds[2],ds[1],ds[0]) 
NSZ={0:4,1:6,2:4,3:1}
def classkey(ds): return tuple(LEV[d] for d in ds)
def class_blocked(dimorder):
    cls={}
    for ds in lex: cls.setdefault(classkey(ds),[]).append(ds)
    out=[]
    for ck in sorted(cls, key=lambda c: c[::-1]):   # classes in order of their level vectors, top position slowest
        sizes=[NSZ[l] for l in ck]
        if dimorder=="lex": perm=[4,3,2,1,0]            # slowest ... fastest
        elif dimorder=="small_fast": perm=sorted(range(5), key=lambda i:(-sizes[i], -i))   # slowest = largest
        elif dimorder=="large_fast": perm=sorted(range(5), key=lambda i:(sizes[i], -i))
        out.extend(sorted(cls[ck], key=lambda ds: tuple(ds[i] for i in perm)))
    return out
for mode in ("lex","small_fast","large_fast"):
    o=class_blocked(mode); assert len(o)==len(T)
    m,t=simulate(o,cap); print("class-blocked", mode, m/t)
