"""Model check (CPU) of the reformulated American-flag pass used by k_bt_sort_mid (bt_flag_pass_fq in
mm2-gb_b200/csrc/backtrack_kernels.cuh): the foreign-token walk + closed-form placement must give exactly the permutation of
the reference's in-place pass (ksort.h:116-139, restated literally in ref_pass) -- including where equal digits end up, which
is what makes the unstable sort's tie order reproducible."""
import random
def ref_pass(keys, digit):
    a=list(keys); n=len(a)
    cnt=[0]*257
    for x in a: cnt[digit(x)+1]+=1
    st=[0]*257
    for d in range(256): st[d+1]=st[d]+cnt[d+1]
    b=st[:256]; e=st[1:257]
    b=list(b)
    k=0
    while k<256:
        if b[k]!=e[k]:
            l=digit(a[b[k]])
            if l!=k:
                tmp=a[b[k]]
                while True:
                    swap=tmp; tmp=a[b[l]]; a[b[l]]=swap; b[l]+=1
                    l=digit(tmp)
                    if l==k: break
                a[b[k]]=tmp; b[k]+=1
            else: b[k]+=1
        else: k+=1
    return a
def fq_pass(keys, digit):
    a=list(keys); m=len(a)
    D=[digit(x) for x in a]
    cnt=[0]*256
    for d in D: cnt[d]+=1
    st=[0]*257
    for d in range(256): st[d+1]=st[d]+cnt[d]
    fst=[0]*256; fen=[0]*256; cur=[0]*256
    tok=[None]*m; fpos=[]; Fd=[]
    g=0
    for r in range(256):
        if st[r+1]>st[r]:
            fst[r]=g; cur[r]=g
            for e in range(st[r],st[r+1]):
                if D[e]!=r:
                    tok[e]=g; fpos.append(e); Fd.append(D[e]); g+=1
                else: tok[e]=-1
            fen[r]=g
    nxt=[None]*g
    ecut=[-1]*256
    hsp=[0]*256
    for k in range(256):
        if st[k+1]>st[k]:
            fe=fen[k]; h=cur[k]; hsp[k]=h
            while h<fe:
                c0=h; h+=1
                carried=c0; d=Fd[c0]
                while d!=k:
                    g2=cur[d]; cur[d]=g2+1; nxt[carried]=('A',g2)
                    carried=g2; d=Fd[g2]
                nxt[carried]=('E',c0)
    for r in range(256):
        if st[r+1]>st[r] and hsp[r]>fst[r]: ecut[r]=fpos[hsp[r]-1]
    out=[None]*m
    for e in range(m):
        d=D[e]
        if tok[e]==-1: dst=e+(1 if e<ecut[d] else 0)
        else:
            kind,v=nxt[tok[e]]
            if kind=='E': dst=fpos[v]
            else: dst=st[d] if v==fst[d] else fpos[v-1]+1
        assert out[dst] is None, (e,dst)
        out[dst]=a[e]
    return out
def test_fq_pass_equals_literal_pass():
  random.seed(1)
  for trial in range(1500):
      n=random.choice([1,2,5,33,100,300,1000])
      nd=random.choice([2,3,5,17,256])
      mode=random.random()
      if mode<0.3: keys=[(random.randrange(nd),i) for i in range(n)]
      elif mode<0.6: # nearly sorted
          keys=sorted([(random.randrange(nd),i) for i in range(n)], key=lambda t:t[0]); 
          for _ in range(n//10+1):
              i=random.randrange(n); keys[i]=(random.randrange(nd),keys[i][1])
      else: # skewed
          keys=[(min(nd-1,int(random.expovariate(1.0))),i) for i in range(n)]
      dg=lambda x:x[0]
      r1=ref_pass(keys,dg); r2=fq_pass(keys,dg)
      assert r1==r2, (trial,n,nd)

