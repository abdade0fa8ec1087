import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
swb = importlib.import_module("smith-waterman_b200")
cols, rows, wpc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
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
t=tr.view(strips,8).cpu().numpy().astype(object)
for s in range(min(strips,6)):
    r=t[s]; n=max(int(r[5]),1); slow=int(r[7])>>32; it=int(r[7])&0xffffffff
    print(f"strip {s}: per group: pre {r[2]/n:7.1f}  steps {r[3]/n:7.1f} ({r[3]/n/8:6.1f}/step)  post {r[4]/n:6.1f} | edge {r[6]/32:6.1f}/step | slow-path {slow} iters {it}")
