import torch, math, time, sys
sys.path.insert(0, "/root/repo")
from diffsci_b200 import ops
DEV="cuda:0"; BF=torch.bfloat16
def run(ndim,B,Cin,Cout,sp):
    torch.manual_seed(0)
    x = torch.randn((B,)+((1,)+sp if ndim==2 else sp)+(Cin,), device=DEV).to(BF)
    dy = torch.randn((B,)+((1,)+sp if ndim==2 else sp)+(Cout,), device=DEV).to(BF)
    _,D,H,W,_ = x.shape
    res = {}
    for wdt in (torch.float32, BF):
        desc = ops.conv_desc(B,D,H,W,Cin,Cout,3,ndim,False,wdt,BF,BF)
        ws = torch.empty(ops.conv_wgrad_ws_bytes(desc), dtype=torch.uint8, device=DEV)
        gw = torch.zeros((Cout,Cin)+(3,)*ndim, device=DEV)
        ops.conv_wgrad(desc,x,dy,gw,ws); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.conv_wgrad(desc,x,dy,gw,ws)
        e1.record(); torch.cuda.synchronize()
        res[wdt]=(gw.clone(), e0.elapsed_time(e1)/5)
    a,b=res[torch.float32][0],res[BF][0]
    err=float((a-b).abs().max()/a.abs().max())
    fl=2.0*B*D*H*W*Cin*Cout*3**ndim
    print(f"wgrad {ndim}d {Cin}->{Cout} {sp} B={B}: tc-vs-ffma max rel {err:.2e}; ffma {res[torch.float32][1]*1e3:.0f} us ({fl/res[torch.float32][1]/1e9:.1f} TF/s), tc {res[BF][1]*1e3:.0f} us ({fl/res[BF][1]/1e9:.1f} TF/s)", flush=True)
    if err > 1e-3:
        # per-tap diagnosis
        d=(a-b).abs().flatten(2).amax(dim=(0,1)); print(" per-tap max abs diff:", [f"{v:.3f}" for v in d.tolist()], "ref absmax", float(a.abs().max()))
run(2,2,64,64,(16,8))
run(2,2,64,64,(32,32))
run(3,1,64,64,(8,8,8))
run(2,5,128,64,(7,7))
run(3,2,64,64,(64,64,64))
run(3,2,128,128,(32,32,32))
run(2,32,64,64,(128,128))
run(2,32,256,256,(32,32))
