import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
import os
cols=rows=int(os.environ.get("SHAPE","8192"))
wpc=int(sys.argv[1]) if len(sys.argv)>1 else 2
SR=int(sys.argv[2]) if len(sys.argv)>2 else 96
dev=torch.device("cuda:0")
a,b=swb.generate(42,cols,rows)
a_d=torch.frombuffer(bytearray(a),dtype=torch.uint8).to(dev); b_d=torch.frombuffer(bytearray(b),dtype=torch.uint8).to(dev)
dH=torch.empty((rows+1)*(cols+1),dtype=torch.int32,device=dev); dP=torch.empty_like(dH)
strips=(rows+SR-1)//SR
for it in range(2):
    tr=torch.zeros(strips*8,dtype=torch.int64,device=dev)
    swb.fill_async(a_d,cols,b_d,rows,dH,dP,cols+1,None,None,warps_per_band=wpc,trace=tr)
    torch.cuda.synchronize()
t=tr.view(strips,8).cpu().numpy().astype(object)
print("strip | steady groups: pre-wait/grp  steps/grp (per step)  post/grp | edge groups 0-3: steps per step, other per group")
for s in list(range(6))+list(range(strips//2,strips//2+4))+[strips-2,strips-1]:
    r=t[s]; n=max(r[5],1)
    drain=(int(r[7])>>32)*16; cons=(int(r[7])&0xffffffff)*16
    print(f"{s:4d} | {r[2]/n:8.1f} {r[3]/n:8.1f} ({r[3]/n/8:6.1f}) {r[4]/n:8.1f} | {r[6]/32:7.1f}  of pre: writers' ring space {drain/n:6.1f}  consumer's ring space {cons/n:6.1f} clk/group")
n=np.maximum(t[:,5],1)
print("mean: pre %.1f steps %.1f (%.1f/step) post %.1f | edge %.1f/step" % ((t[:,2]/n).mean(), (t[:,3]/n).mean(), (t[:,3]/n).mean()/8, (t[:,4]/n).mean(), (t[:,6]/32).mean()))
