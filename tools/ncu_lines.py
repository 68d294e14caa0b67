import csv,sys,subprocess,collections
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum']
for i,h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h):
        try:
            v=float(vals[i].replace(',',''))
        except: continue
        if 'issue_stalled' in h and v<0.3: continue
        print("%-80s %-10s %s"%(h,units[i],vals[i]))
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur_file=None;cur_line=None;hdr=None
agg=collections.defaultdict(lambda:[0,0])
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if r[0]=='Function Name': continue
    if r[0].isdigit(): cur_line=int(r[0]); continue
    if r[0]=='' and len(r)>8 and r[2].startswith('0x'):
        tail=r[-(len(hdr)-4):]
        try: s=float(tail[2]); n=float(tail[3])
        except: continue
        a=agg[(cur_file,cur_line)]; a[0]+=s; a[1]+=n
tot=sum(a[0] for a in agg.values()); totn=sum(a[1] for a in agg.values())
print("samples",tot,"inst",totn)
src={}
import os
for fn in ('ob_decode_pipe.cu','ob_decode_tile.cuh','ob_ptx.cuh','ob_cloud.cu'):
    p='/root/repo/ouster-sdk_b200/csrc/'+fn
    if os.path.exists(p): src[fn]=open(p).read().split('\n')
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for (fn,ln),a in sorted(agg.items(),key=lambda kv:-kv[1][0])[:N]:
    line=src[fn][ln-1].strip()[:90] if fn in src and ln-1<len(src[fn]) else ''
    print("%-20s %4d  samp %5.1f%%  inst %5.1f%% (%.2fM) %s"%(fn,ln,100*a[0]/tot,100*a[1]/totn,a[1]/1e6,line))
# by file+range buckets of 25 lines
b=collections.defaultdict(lambda:[0,0])
for (fn,ln),a in agg.items():
    k=(fn,ln//20*20); b[k][0]+=a[0]; b[k][1]+=a[1]
print("---- buckets")
for k,a in sorted(b.items(),key=lambda kv:-kv[1][1])[:25]:
    print("%-22s %4d-%4d samp %5.1f%% inst %5.1f%% (%.2fM)"%(k[0],k[1],k[1]+19,100*a[0]/tot,100*a[1]/totn,a[1]/1e6))
