"""Column bands per row block of the packed operators (PMC_DUMP_SELL dumps, gpurun_out/sell/): how many rows of the gathered
vector and of the weight vector a block of 224 / 448 rows touches, and in how many contiguous segments -- the sizing of a
shared-memory x-band staging."""
import numpy as np, struct, sys, glob
def load(path):
    b=open(path,'rb').read()
    rows,cols,weighted,noff=struct.unpack('4i',b[:16]); nb,=struct.unpack('q',b[16:24])
    off=np.frombuffer(b[24:24+4*noff],dtype=np.int32); pk=b[24+4*noff:]
    es=16 if weighted else 12
    S=16
    ent=[]  # per slice: (vals[w,16], cols[w,16], widx[w,16] or None)
    for sl in range(noff-1):
        w=off[sl+1]-off[sl]; base=off[sl]*S*es
        v=np.frombuffer(pk[base:base+w*S*8],dtype=np.float64).reshape(w,S)
        c=np.frombuffer(pk[base+w*S*8:base+w*S*12],dtype=np.int32).reshape(w,S)
        wi=np.frombuffer(pk[base+w*S*12:base+w*S*16],dtype=np.int32).reshape(w,S) if weighted else None
        ent.append((v,c,wi))
    return rows,cols,weighted,ent
def analyze(path, bsl, Nf=None):
    rows,cols,weighted,ent=load(path)
    nsl=len(ent)
    tot_x=[];tot_w=[]; nseg=[]
    for b0 in range(0,nsl,bsl):
        cs=[];ws=[]
        for sl in range(b0,min(nsl,b0+bsl)):
            v,c,wi=ent[sl]
            m=v!=0
            cs.append(c[m]);
            if weighted: ws.append(wi[m])
        cs=np.unique(np.concatenate(cs))
        # split into segments with gap > 64
        gaps=np.where(np.diff(cs)>64)[0]
        segs=np.split(cs,gaps+1)
        lenx=sum(s[-1]-s[0]+1 for s in segs)
        tot_x.append(lenx); nseg.append(len(segs))
        if weighted:
            wsu=np.unique(np.concatenate(ws))
            gaps=np.where(np.diff(wsu)>64)[0]
            segw=np.split(wsu,gaps+1)
            tot_w.append(sum(s[-1]-s[0]+1 for s in segw))
    print(path.split('/')[-1], "rows",rows,"blocks of",bsl*16,"rows: x band rows max/mean",max(tot_x),np.mean(tot_x),"segments max",max(nseg), ("w band max/mean %d %.0f"%(max(tot_w),np.mean(tot_w)) if weighted else ""))
for f in ["op_009_r13056_c17152_w","op_010_r4096_c13056_p","op_006_r4096_c13056_p","op_007_r13056_c17152_p","op_016_r4096_c4096_w","op_008_r4096_c4096_p","op_033_r1728_c2240_w"]:
    for bsl in (14,28):
        analyze("/root/repo/gpurun_out/sell/%s.bin"%f, bsl)
