import csv, sys, collections
f=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(f)))
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[0]!="Address"]
tot_inst=sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
tot_samp=sum(float(r[ix["# Samples"]] or 0) for r in data)
print("total warp-instr", tot_inst, "samples", tot_samp, "sass lines", len(data))
# opcode histogram
hist=collections.Counter(); sh=collections.Counter()
for r in data:
    src=r[ix["Source"]].strip()
    op=src.split()[0] if not src.startswith('@') else src.split()[1]
    op=op.split('.')[0]
    hist[op]+=float(r[ix["Instructions Executed"]] or 0)
    sh[op]+=float(r[ix["# Samples"]] or 0)
print("opcode: %instr  %samples")
for op,c in hist.most_common(28):
    print(f"  {op:10s} {100*c/tot_inst:6.2f}  {100*sh[op]/tot_samp:6.2f}")
# stall reasons total
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot=collections.Counter()
for r in data:
    for s in stalls:
        v=r[ix[s]]
        if v: tot[s]+=float(v)
ss=sum(tot.values())
print("stalls:", {k: round(100*v/ss,1) for k,v in tot.most_common(10)})
print("top lines by samples:")
order=sorted(range(len(data)), key=lambda i:-float(data[i][ix["# Samples"]] or 0))[:topn]
for i in sorted(order):
    r=data[i]
    st={s: float(r[ix[s]] or 0) for s in stalls}
    top=max(st,key=st.get)
    print(f"  {i:5d} {100*float(r[ix['# Samples']] or 0)/tot_samp:5.2f}% inst {float(r[ix['Instructions Executed']] or 0)/1e6:7.2f}M thr {r[ix['Avg. Threads Executed']]:>5s} {top:18s} {r[ix['Source']][:90]}")
