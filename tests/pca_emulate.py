"""Ad-hoc (not a test, CPU only): numpy model of the subspace iteration of csrc/pca.cu (random block, Cholesky QR, Rayleigh-Ritz,
Chebyshev filter with the same bounds and degree rule), to count operator applications / CholQR passes / Rayleigh-Ritz steps of
variants before spending GPU time on them.  `cost()` prices the counts with the per-launch times measured at N = 2000.
   python tests/pca_emulate.py [N]
Findings, round 1 (N = 600..4000, seeds 1-5): the model reproduces the GPU's counts at N = 2000 (32 applications, 3 Rayleigh-
Ritz steps, 3 iterations).  Replacing the start Rayleigh-Ritz step by Rayleigh quotients + a 1-norm bound for the top
(start='norm1') saves one eigensolve and ~20 % of the modelled PCA time at N = 2000 / 3000, but costs one to two extra
iterations at N = 600 / 1000 / 1200 / 4000: not a robust default.  inner = 4, cond cap 1e5, b = 256 stay the best all-round
for N >= 1500.  Below that (Nf = 600..1200, where the 256-wide block is a quarter to a half of the spectrum) the degree rule
leaves most filter rounds at degree 1 and the solve takes 4-6 iterations; a cond cap of 1e7 brings it back to 3 in the model
(21 -> 11 modelled ms at N = 600), but one-pass Cholesky QR at that conditioning has to be checked on the GPU first."""
import sys, os, numpy as np, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tadpole_oracle as O
from tadpole_b200.synth import synth_hic

def setup(n, seed=1):
    m = synth_hic(n, seed=seed)
    lm = O.load_mat_numeric(m)
    C = O.sparse_cor(lm.mat)
    Xc = C - C.mean(axis=0, keepdims=True)
    M = Xc @ Xc.T
    return M

def cholqr(Y, passes, cnt):
    for _ in range(passes):
        G = Y.T @ Y
        try:
            L = np.linalg.cholesky(G)
            Y = np.linalg.solve(L, Y.T).T
        except np.linalg.LinAlgError:
            Y, _ = np.linalg.qr(Y); cnt['fail'] = cnt.get('fail', 0) + 1
        cnt['chol'] += 1
    return Y

def rr(M, Y, cnt):
    W = M @ Y; cnt['apps'] += 1
    T = Y.T @ W; T = 0.5*(T+T.T)
    w, Z = np.linalg.eigh(T)
    idx = np.argsort(-w); w = w[idx]; Z = Z[:, idx]
    cnt['rr'] += 1
    return Y @ Z, w

def solve(M, k=200, b=256, inner=4, tol=1e-12, start='rr', maxdeg=24, condcap=1e5, seed=0, verbose=False):
    n = M.shape[0]
    rng = np.random.default_rng(seed)
    cnt = dict(apps=0, chol=0, rr=0, resid=0)
    Y = rng.uniform(-1, 1, size=(n, b))
    Y = cholqr(Y, 2, cnt)
    if start == 'rr':
        Y, theta = rr(M, Y, cnt)
    else:
        # heuristic start: Rayleigh quotients only (no eigensolve): top from a norm bound, cut from the mean eigenvalue
        W = M @ Y; cnt['apps'] += 1
        rq = np.sort(np.einsum('ij,ij->j', Y, W))[::-1]
        theta = rq.copy()
        theta[0] = np.abs(M).sum(axis=1).max() if start == 'norm1' else rq[0]
    last = 1.0
    for it in range(1, 40):
        top, thk = theta[0], theta[k-1]
        cut = theta[b-1]
        if not (cut > 0): 
            pos = theta[k:][theta[k:] > 0]; cut = pos[-1] if pos.size else 0.0
        if not (cut > 0) or not (thk > cut): cut = 0.5*thk if thk > 0 else 1e-300
        e = c = 0.5*cut
        sig1 = e/(top-c)
        xk, x1 = (thk-c)/e, (top-c)/e
        deg = 1
        for md in range(2, maxdeg+1):
            ratio = math.cosh(md*math.acosh(x1))/math.cosh(md*math.acosh(max(xk,1.0)))
            if ratio <= condcap: deg = md
            else: break
        F1 = (sig1/e)*(M@Y - c*Y); cnt['apps'] += 1
        MY = (e/sig1)*F1 + c*Y
        res = np.linalg.norm(MY[:, :k] - Y[:, :k]*theta[:k], axis=0).max()/top
        cnt['resid'] += 1
        if verbose: print(f"it={it} res={res:.2e} top={top:.3e} thk={thk:.3e} cut={cut:.3e} deg={deg} apps={cnt['apps']}")
        if res <= tol: return cnt, it, True
        for r in range(inner):
            if r > 0:
                F1 = (sig1/e)*(M@Y - c*Y); cnt['apps'] += 1
            sig = sig1; P0, P1 = Y, F1
            for j in range(2, deg+1):
                sig2 = 1.0/(2.0/sig1 - sig)
                P2 = 2*(sig2/e)*(M@P1 - c*P1) - sig*sig2*P0; cnt['apps'] += 1
                P0, P1 = P1, P2; sig = sig2
            Y = cholqr(P1, 2 if r+1 == inner else 1, cnt)
        Y, theta = rr(M, Y, cnt)
    return cnt, it, False

def cost(cnt):   # ms at N = 2000 from the measured per-launch times
    return cnt['apps']*0.075 + cnt['chol']*0.205 + cnt['rr']*2.1 + cnt['resid']*0.14

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    M = setup(n)
    print("n", M.shape)
    for kw in [dict(), dict(start='rq'), dict(start='norm1'), dict(inner=3), dict(inner=5), dict(b=224), dict(b=288), dict(b=320),
               dict(condcap=1e6), dict(condcap=1e7), dict(inner=3, condcap=1e6), dict(inner=2, condcap=1e7)]:
        t=time.time(); cnt, it, ok = solve(M, **kw)
        print(kw, cnt, "iters", it, "ok", ok, "model ms %.2f" % cost(cnt), "(%.1fs)" % (time.time()-t), flush=True)
