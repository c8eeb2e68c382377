import torch, time
dev='cuda'
for D in (512,1024):
    g=torch.Generator().manual_seed(0)
    X=torch.randn(300,D,generator=g).to(dev)*0.05
    A=(X.t()@X)/300*0.5+2e-4*torch.eye(D,device=dev)
    A=(A+A.t())/2
    I=torch.eye(D,device=dev)
    def t(fn,n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): r=fn()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n*1e3, r
    ref=torch.linalg.inv(A.double())
    for name,fn in [("linalg.inv", lambda: torch.linalg.inv(A)), ("cholesky+cholesky_inverse", lambda: torch.cholesky_inverse(torch.linalg.cholesky(A))),
                    ("cholesky+cholesky_solve(I)", lambda: torch.cholesky_solve(I, torch.linalg.cholesky(A))), ("linalg.solve(A,I)", lambda: torch.linalg.solve(A,I))]:
        us,r=t(fn)
        err=float((r.double()-ref).abs().max()/ref.abs().max())
        print(f"D={D} {name:28s} {us:9.1f} us  rel err vs fp64 {err:.2e}  cond {float(torch.linalg.cond(A.double())):.1e}")
