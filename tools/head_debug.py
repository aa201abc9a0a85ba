import numpy as np, torch, sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, os.getcwd())
from tests.test_head import _case, _rel
from oracle import head_f64
from aes_lac_2018_b200.head import _HeadFn
for (T,B,H,V,shift) in [(50,4,800,29,0.0),(7,1,16,5,0.0),(20,2,64,64,0.0),(300,8,800,29,3.0)]:
    N=T*B
    x, W, g, b, rm, rv, dl = _case(T*1000+V, N, H, V, shift)
    o, cache, nrm, nrv = head_f64.head_forward(x, W, g, b, rm, rv, True)
    dx, dW, dg, db = head_f64.head_backward(dl, cache)
    dev="cuda"
    xt = torch.tensor(x, device=dev).view(T,B,H).requires_grad_(True)
    Wt = torch.tensor(W, device=dev).requires_grad_(True)
    gt = torch.tensor(g, device=dev).requires_grad_(True)
    bt = torch.tensor(b, device=dev).requires_grad_(True)
    out = _HeadFn.apply(xt, Wt, gt, bt, torch.tensor(rm, device=dev), torch.tensor(rv, device=dev), True, 1e-5, 0.1, False)
    out.backward(torch.tensor(dl, device=dev).view(T,B,V))
    gdx = xt.grad.cpu().numpy().reshape(N,H); gdW = Wt.grad.cpu().numpy()
    print((T,B,H,V), "out", _rel(out.detach().cpu().numpy().reshape(N,V), o), "dx", _rel(gdx, dx), "dW", _rel(gdW, dW), "dg", _rel(gt.grad.cpu().numpy(), dg), "db", _rel(bt.grad.cpu().numpy(), db))
    e = np.abs(gdW - dW); print("  dW err by v (max over h):", np.round(e.max(1)/np.abs(dW).max(), 4)[:8], " by h-block:", [float(np.round(e[:, i:i+4].max()/np.abs(dW).max(),4)) for i in range(0, min(H,32), 4)])
    e = np.abs(gdx - dx); print("  dx err by row:", np.round(e.max(1)/np.abs(dx).max(), 4)[:6], "by h:", np.round(e.max(0)/np.abs(dx).max(), 4)[:8])
    if (T,B,H,V)==(7,1,16,5):
        print(np.round(gdW[:, :8],4)); print(np.round(dW[:, :8],4))
