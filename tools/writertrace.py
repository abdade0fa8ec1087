import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
import os
cols=rows=int(os.environ.get("SHAPE","8192"))
wpc=int(sys.argv[1]) if len(sys.argv)>1 else 2
SR=64
dev=torch.device("cuda:0")
a,b=swb.generate(42,cols,rows)
a_d=torch.frombuffer(bytearray(a),dtype=torch.uint8).to(dev); b_d=torch.frombuffer(bytearray(b),dtype=torch.uint8).to(dev)
dH=torch.empty((rows+1)*(cols+1),dtype=torch.int32,device=dev); dP=torch.empty_like(dH)
strips=(rows+SR-1)//SR
for it in range(2):
    tr=torch.zeros(strips*8,dtype=torch.int64,device=dev)
    swb.fill_async(a_d,cols,b_d,rows,dH,dP,cols+1,None,None,warps_per_band=wpc,trace=tr)
    torch.cuda.synchronize()
t=tr.view(strips,8).cpu().numpy().astype(float)
n=np.maximum(t[:,7],1)
print("writer 0 of each strip, interior rounds: wait clk/round, work clk/round")
for s in list(range(4))+[strips//2, strips//2+1, strips-2, strips-1]:
    print(s, round(t[s,5]/n[s],1), round(t[s,6]/n[s],1), int(n[s]))
print("mean wait %.1f work %.1f" % ((t[:,5]/n).mean(), (t[:,6]/n).mean()))
