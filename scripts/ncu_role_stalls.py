import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
# split by kernel
starts=[k for k,r in enumerate(rows) if r and r[0]=="Kernel Name"]
starts.append(len(rows))
for si in range(len(starts)-1):
    blk=rows[starts[si]:starts[si+1]]
    print("=====",blk[0][1][:70])
    hdr=blk[1]; data=blk[2:]
    isrc=hdr.index("Source"); isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed")
    stalls=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    alloc=[k for k,r in enumerate(data) if "USETMAXREG" in r[isrc]]
    def agg(lo,hi,name):
        tot=0; st={}; ex=0; fp64=0; lds=0; loc=0
        for r in data[lo:hi]:
            n=int(r[isamp] or 0); tot+=n
            e=int(r[iex] or 0); ex+=e
            op=r[isrc].split()[0] if r[isrc].split() else ""
            if op.startswith("@"): op=r[isrc].split()[1]
            if op[:4] in ("DFMA","DADD","DMUL","DSET"): fp64+=e
            if op.startswith("LDS") or op.startswith("STS"): lds+=e
            if op.startswith("LDL") or op.startswith("STL"): loc+=e
            for i in stalls:
                v=int(r[i] or 0); st[hdr[i]]=st.get(hdr[i],0)+v
        print(name,"samples",tot,"inst",ex,"fp64",fp64,"lds/sts",lds,"local",loc)
        for k,v in sorted(st.items(),key=lambda kv:-kv[1])[:8]:
            print("    ",k,v,round(100*v/max(tot,1),1))
    if len(alloc)==2:
        a,b=alloc
        agg(0,a,"prologue"); agg(a,b,"face warps"); agg(b,len(data),"helper warps")
    else:
        agg(0,len(data),"all")
