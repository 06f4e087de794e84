# prototype of the multi-pass NTT index scheme (host-side math check, python ints)
import random
P=2013265921
def gen(bits):
    g=0x1a427a41
    for _ in range(27-bits): g=g*g%P
    return g
def naive_dft(c,N):
    w=gen(N.bit_length()-1)
    return [sum(c[j]*pow(w,i*j,P) for j in range(len(c)))%P for i in range(N)]
def brev(x,bits):
    r=0
    for b in range(bits): r|=((x>>b)&1)<<(bits-1-b)
    return r
def dif_inplace(a,R,wR):
    # natural in -> bit reversed out, DIF radix 2; a list len R
    r=R.bit_length()-1
    for t in range(r):
        half=R>>(t+1)
        for b in range(R//2):
            g,pos=divmod(b,half)
            i0=g*2*half+pos;i1=i0+half
            u,v=a[i0],a[i1]
            a[i0]=(u+v)%P
            a[i1]=(u-v)*pow(wR,pos<<t,P)%P
def multipass(x,H,N,rs):
    n=N.bit_length()-1
    assert sum(rs)==n
    wN=gen(n)
    buf=[0]*N
    # pass i: data viewed [outer][R][S]
    cur=None
    consumed=0
    for i,r in enumerate(rs):
        R=1<<r
        S=N>>(consumed+r)
        M=R*S
        outer=N//M
        wR=gen(r); wM=gen(M.bit_length()-1)
        last=(i==len(rs)-1)
        new=[0]*N
        for o in range(outer):
            for s in range(S):
                if i==0:
                    col=[x[rr*S+s] if rr*S+s<H else 0 for rr in range(R)]
                else:
                    col=[cur[o*M+rr*S+s] for rr in range(R)]
                dif_inplace(col,R,wR)
                for k in range(R):
                    v=col[brev(k,r)]*pow(wM,s*k,P)%P
                    if not last:
                        new[o*M+k*S+s]=v
                    else:
                        # o encodes (k_1..k_{m-1}) in in-place mixed radix order (k_1 most significant)
                        # natural: k = k_1 + T_2 k_2 + ... ; T_m = N/R
                        # convert o digits
                        digs=[];oo=o
                        for rr in reversed(rs[:-1]):
                            digs.append(oo&((1<<rr)-1));oo>>=rr
                        digs=digs[::-1]  # k_1..k_{m-1}
                        q=0;T=1
                        for d,rr in zip(digs,rs[:-1]):
                            q+=d*T;T<<=rr
                        new[q+T*k]=v
        cur=new
        consumed+=r
    return cur
random.seed(1)
for (H,N,rs) in [(8,16,[2,2]),(16,32,[2,3]),(32,64,[2,2,2]),(16,64,[3,1,2]),(64,64,[3,3]),(8,8,[3]),(4,8,[3]),(16,64,[6]),(32,128,[2,3,2])]:
    x=[random.randrange(P) for _ in range(H)]
    assert multipass(x,H,N,rs)==naive_dft(x,N),(H,N,rs)
print("ok")
