"""Offline: PGS of one heavy-env record in velocity form (u) and impulse form (per-row sums), float32 vs float64."""
import numpy as np, sys
HDR, ROW, N, NT = 84, 36, 9, 45
MI, MRHS, LRHS, LIM, GEAR, NC = 0, 45, 54, 63, 65, 83
F1, F2, GR = 7, 8, -1.0
def tri(i, j): return i*(i+1)//2 + j if i >= j else j*(j+1)//2 + i
def rows_of(rec, dt):
    nc = int(rec[NC])
    Mi = np.array([[rec[MI + tri(i, j)] for j in range(N)] for i in range(N)], dt)
    J, V, rhs, dinv, cfm, lo, hi, kind = [], [], [], [], [], [], [], []
    hi_arm, hi_fin, hi_gear = 100000/60., 1000/60., 50/60.
    for i in range(N):
        e = np.zeros(15, dt); e[i] = 1; v = np.zeros(15, dt); v[:N] = Mi[:, i]
        J.append(e); V.append(v); rhs.append(rec[MRHS+i]); dinv.append(dt(1)/Mi[i, i]); cfm.append(0); hi.append(hi_arm if i < 7 else hi_fin); lo.append(-hi[-1]); kind.append('m')
    e = np.zeros(15, dt); e[F1] = 1; e[F2] = GR; v = np.zeros(15, dt); v[:N] = Mi[:, F1] + dt(GR)*Mi[:, F2]
    J.append(e); V.append(v); rhs.append(rec[GEAR]); dinv.append(rec[GEAR+1]); cfm.append(0); hi.append(hi_gear); lo.append(-hi_gear); kind.append('g')
    mu = []
    for c in range(nc):
        for k in range(3):
            r = rec[HDR + (3*c+k)*ROW: HDR + (3*c+k+1)*ROW]
            J.append(r[0:15].astype(dt)); V.append(r[16:31].astype(dt)); rhs.append(r[32]); dinv.append(r[33]); cfm.append(r[34] if k == 0 else 0); lo.append(0); hi.append(1e10); kind.append('nab'[k])
        mu.append(rec[HDR + 3*c*ROW + 35])
    return nc, np.array(J, dt), np.array(V, dt), np.array(rhs, dt), np.array(dinv, dt), np.array(cfm, dt), np.array(lo, dt), np.array(hi, dt), np.array(mu, dt)
def solve(rec, dt, form, sweeps=50):
    nc, J, V, rhs, dinv, cfm, lo, hi, mu = rows_of(rec, dt)
    R = len(J); lam = np.zeros(R, dt); u = np.zeros(15, dt)
    if form == 'impulse':
        A = (J @ V.T).astype(dt); s = np.zeros(R, dt)
    def S(r): return (J[r] @ u).astype(dt) if form == 'velocity' else s[r]
    def apply(r, d):
        nonlocal u, s
        if form == 'velocity': u = (u + V[r]*d).astype(dt)
        else: s = (s + A[:, r]*d).astype(dt)
    for it in range(sweeps):
        order = list(range(0, 9)) + [9] if it & 1 else [9] + list(range(8, -1, -1))
        for r in order:
            d = rhs[r] - S(r)*dinv[r]; sm = lam[r] + d; sc = min(max(sm, lo[r]), hi[r]); d = d if sc == sm else sc - lam[r]
            lam[r] = dt(lam[r] + d); apply(r, dt(d))
        for c in range(nc):
            r = 10 + 3*c
            d = rhs[r] - lam[r]*cfm[r] - S(r)*dinv[r]; sm = lam[r] + d; sc = max(sm, dt(0)); d = d if sc == sm else sc - lam[r]
            lam[r] = dt(lam[r] + d); apply(r, dt(d))
        for c in range(nc):
            ra, rb = 11 + 3*c, 12 + 3*c
            lim = mu[c]*lam[10+3*c]
            da = rhs[ra] - S(ra)*dinv[ra]; db = rhs[rb] - S(rb)*dinv[rb]
            sa, sb = lam[ra] + da, lam[rb] + db
            l2 = sa*sa + sb*sb
            if l2 > lim*lim:
                ln = np.sqrt(l2)
                if ln > lim:
                    sc = lim/ln if ln > 0 else 0; sa *= sc; sb *= sc; da = sa - lam[ra]; db = sb - lam[rb]
            lam[ra] = dt(lam[ra] + da); lam[rb] = dt(lam[rb] + db); apply(ra, dt(da)); apply(rb, dt(db))
    return (V.T @ lam).astype(np.float64), lam.astype(np.float64)
def solve2(rec, dt, sweeps=50):
    """the short-chain update of heavy_solve_dela: x = (app + rhs - app cfm) - s dinv, app' = clamp(x), delta = app' - app"""
    nc, J, V, rhs, dinv, cfm, lo, hi, mu = rows_of(rec, dt)
    R = len(J); lam = np.zeros(R, dt); A = (J @ V.T).astype(dt); s = np.zeros(R, dt)
    f = dt
    for it in range(sweeps):
        order = list(range(0, 9)) + [9] if it & 1 else [9] + list(range(8, -1, -1))
        for r in order:
            c = f(lam[r] + rhs[r]); x = f(c - f(s[r]*dinv[r])); xn = min(max(x, lo[r]), hi[r]); d = f(xn - lam[r]); lam[r] = xn; s = (s + A[:, r]*d).astype(dt)
        for c_ in range(nc):
            r = 10 + 3*c_
            c = f(f(lam[r] + rhs[r]) - f(lam[r]*cfm[r])); x = f(c - f(s[r]*dinv[r])); xn = min(max(x, f(0)), f(1e10)); d = f(xn - lam[r]); lam[r] = xn; s = (s + A[:, r]*d).astype(dt)
        for c_ in range(nc):
            ra, rb = 11 + 3*c_, 12 + 3*c_
            lim = f(mu[c_]*lam[10+3*c_])
            xa = f(f(lam[ra] + rhs[ra]) - f(s[ra]*dinv[ra])); xb = f(f(lam[rb] + rhs[rb]) - f(s[rb]*dinv[rb]))
            l2 = f(xa*xa + xb*xb)
            if l2 > lim*lim:
                sc = f(lim * f(1/np.sqrt(l2))); xa = f(xa*sc); xb = f(xb*sc)
            da = f(xa - lam[ra]); db = f(xb - lam[rb]); lam[ra] = xa; lam[rb] = xb
            s = (s + A[:, ra]*da).astype(dt); s = (s + A[:, rb]*db).astype(dt)
    return (V.T @ lam).astype(np.float64), lam.astype(np.float64)
if __name__ == "__main__":
    d = np.load("gpurun_out/hrec_dump.npz")
    rng = np.random.default_rng(0)
    for key in ["step0", "step2", "step5"]:
        recs = d[key]
        ok = [i for i in range(len(recs)) if 1 <= recs[i][NC] <= 16 and np.isfinite(recs[i]).all() and recs[i][MI] > 0]
        pick = ok[:8]
        for i in pick:
            rec = recs[i]
            u64, l64 = solve(rec, np.float64, 'velocity')
            u64i, _ = solve(rec, np.float64, 'impulse')
            u32v, _ = solve(rec, np.float32, 'velocity')
            u32i, _ = solve2(rec, np.float32)
            u64_100, _ = solve(rec, np.float64, 'velocity', 100)
            sc = np.abs(u64).max()
            print(f"{key} slot {i} nc={int(rec[NC])} |u|max={sc:.3g} lam_max={np.abs(l64).max():.3g}  f64 imp-vs-vel {np.abs(u64i-u64).max():.2e}  f32vel err {np.abs(u32v-u64).max():.2e}  f32imp err {np.abs(u32i-u64).max():.2e}   50 vs 100 sweeps {np.abs(u64_100-u64).max():.2e}")
