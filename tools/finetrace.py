import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
cols=rows=4096
dev=torch.device("cuda:0")
a,b=swb.generate(42,cols,rows)
a_d=torch.frombuffer(bytearray(a),dtype=torch.uint8).to(dev); b_d=torch.frombuffer(bytearray(b),dtype=torch.uint8).to(dev)
dH=torch.empty((rows+1)*(cols+1),dtype=torch.int32,device=dev); dP=torch.empty_like(dH)
strips=(rows+31)//32
for it in range(2):
    tr=torch.zeros(strips*8,dtype=torch.int64,device=dev)
    swb.fill_async(a_d,cols,b_d,rows,dH,dP,cols+1,None,None,warps_per_band=2,trace=tr)
    torch.cuda.synchronize()
t=tr.view(strips,8).cpu().numpy().astype("float64"); t=(t-t[:,0].min())/1000
print("strip  enter   gate     g1     g2     g3     g4     g5     g8")
for s in list(range(8))+[60,61,62,63]:
    print(f"{s:4d} "+" ".join(f"{x:7.1f}" for x in t[s]))
