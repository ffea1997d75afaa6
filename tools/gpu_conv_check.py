"""GPU diagnostic: runs every conv-op case, prints error norms, keeps going after failures."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tests.test_conv_ops_gpu as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
m = T.load()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
sel = [c for c in T.CASES if which in ("all", c[0])]
for case in sel:
    name, tr, ks, st, c0, c1, co, dims = case
    try:
        x0, x1, w, b = T.make(case)
        y_ref = T.ref_forward(case, x0, x1, w.half().float(), b).cpu().numpy()
        t0 = time.time()
        y, stats = m.conv_forward(x0.numpy(), w.numpy(), b.numpy(), None if x1 is None else x1.numpy(),
                                  transposed=bool(tr), ks=ks, stride=st, want_stats=True)
        e = T.rel(y, y_ref)
        print(f"FWD {name:22s} rel={e:.3e} nan={np.isnan(y).sum()} maxabs={np.abs(y - y_ref).max():.3e} t={time.time()-t0:.2f}s", flush=True)
        if not tr:
            v = y_ref.reshape(co, -1).astype(np.float64)
            print(f"    stats sum err={np.abs(stats[0]-v.sum(1)).max():.3e} sumsq rel={np.abs(stats[1]/(v*v).sum(1)-1).max():.3e}", flush=True)
    except Exception as ex:
        print(f"FWD {name} EXC {ex}", flush=True)
        if "failed" in str(ex) or "timeout" in str(ex):
            print("context dead; stopping"); sys.exit(3)
for case in sel:
    name, tr, ks, st, c0, c1, co, dims = case
    try:
        import torch.nn.functional as F
        x0, x1, w, b = T.make(case, 1)
        x = (x0 if x1 is None else torch.cat([x0, x1], 0))[None].cuda().requires_grad_(True)
        wq = w.half().float().cuda().requires_grad_(True)
        y = F.conv_transpose3d(x, wq, None, stride=2) if tr else F.conv3d(x, wq, None, stride=st, padding=(ks - 1) // 2)
        g = torch.Generator().manual_seed(7)
        dy = torch.randn(y.shape[1:], generator=g).half().float()
        y.backward(dy[None].cuda())
        gx_ref = x.grad[0].cpu().numpy(); gw_ref = wq.grad.cpu().numpy()
        gx0, gx1, gw = m.conv_backward(x0.numpy(), w.numpy(), dy.numpy(), None if x1 is None else x1.numpy(),
                                       transposed=bool(tr), ks=ks, stride=st)
        msg = f"BWD {name:22s} gx0={T.rel(gx0, gx_ref[:c0]):.3e}"
        if c1: msg += f" gx1={T.rel(gx1, gx_ref[c0:]):.3e}"
        msg += f" gw={T.rel(gw, gw_ref):.3e} nan={np.isnan(gx0).sum()}/{np.isnan(gw).sum()}"
        print(msg, flush=True)
    except Exception as ex:
        print(f"BWD {name} EXC {ex}", flush=True)
        if "failed" in str(ex) or "timeout" in str(ex):
            print("context dead; stopping"); sys.exit(3)
