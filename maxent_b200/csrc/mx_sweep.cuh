// Fused MaxEnt alpha sweep for sm_100a (B200).
//
// One persistent CTA per SM marches T spectra through the whole alpha mesh in lock-step "rounds".
// A round evaluates the cost function Q(v - dv) for one trial vector per spectrum:
//     x = V' t          (DMMA, spectra are the M dimension of the 8x8x4 f64 MMA)
//     H = D exp(x)      (v -> H -> A exponential map, entropy terms)      functions.py:739-741,508-510
//     y = V'^T H        (DMMA)        chi2 = sum_i (xi_i y_i - g~_i)^2 + c0           functions.py:358-360
// and, for spectra that start a new Levenberg iteration, the Hessian pieces
//     Z = V'^T diag(w) V'   (DMMA)    f = Z (g + alpha v),  J = eta Z Lambda Z + alpha Z
//                                                      maxent_cost_function.py:95-165 in singular space
// Between rounds one owner warp per spectrum advances that spectrum's Levenberg-Marquardt state
// machine (levenberg_minimizer.py:123-248, mirrored branch for branch) and solves
// (J + mu I) dv = f with a packed Cholesky in shared memory; a failed factorisation counts as
// Q = NaN ("raise mu"), exactly the role garbage LU solutions play in the reference.
// V' is streamed L2 -> shared memory with cp.async in a swizzled 8x8-tile layout (mx_layout_V)
// that makes every DMMA fragment load bank-conflict free.
#pragma once
#include "mx_common.cuh"

namespace mx {

// ------------------------------------------------------------------------------------------
// packed lower-triangular helpers (row-major: (i,j), i>=j at i(i+1)/2 + j)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int pidx(int i, int j) { return (i * (i + 1)) / 2 + j; }
__device__ __forceinline__ int sidx(int a, int b) { return a >= b ? pidx(a, b) : pidx(b, a); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// In-place Cholesky of the packed matrix L (s x s) by one warp.  invd[j] = 1/L_jj.
// Returns false (warp-uniform) on a non-positive / NaN pivot.
static __device__ bool chol_packed(double* __restrict__ L, int s, double* __restrict__ invd, int lane) {
    for (int j = 0; j < s; ++j) {
        const double* rowj = L + pidx(j, 0);
        const int i0 = j + lane, i1 = j + lane + 32;
        const bool a0 = i0 < s, a1 = i1 < s;
        const double* r0 = L + pidx(a0 ? i0 : j, 0);
        const double* r1 = L + pidx(a1 ? i1 : j, 0);
        double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
        int k = 0;
        for (; k + 1 < j; k += 2) {
            const double b0 = rowj[k], b1 = rowj[k + 1];
            s00 = fma(r0[k], b0, s00); s01 = fma(r0[k + 1], b1, s01);
            s10 = fma(r1[k], b0, s10); s11 = fma(r1[k + 1], b1, s11);
        }
        if (k < j) { const double b0 = rowj[k]; s00 = fma(r0[k], b0, s00); s10 = fma(r1[k], b0, s10); }
        const double v0 = r0[j] - (s00 + s01);
        const double v1 = a1 ? (L[pidx(i1, j)] - (s10 + s11)) : 0.0;
        const double d = __shfl_sync(0xffffffffu, v0, 0);
        if (!(d > 0.0)) return false;
        const double r = sqrt(d);
        const double inv = 1.0 / r;
        __syncwarp();
        if (lane == 0) { L[pidx(j, j)] = r; invd[j] = inv; }
        else if (a0) L[pidx(i0, j)] = v0 * inv;
        if (a1) L[pidx(i1, j)] = v1 * inv;
        __syncwarp();
    }
    return true;
}

// Solve L L^T x = b with the factor from chol_packed.  b and x are length-s shared vectors.
static __device__ void chol_solve(const double* __restrict__ L, const double* __restrict__ invd, int s,
                           const double* __restrict__ b, double* __restrict__ x, int lane) {
    double z0 = lane < s ? b[lane] : 0.0;
    double z1 = lane + 32 < s ? b[lane + 32] : 0.0;
    // forward: L z = b (column oriented)
    for (int j = 0; j < s; ++j) {
        const double src = j < 32 ? z0 : z1;
        const double zj = __shfl_sync(0xffffffffu, src, j & 31) * invd[j];
        if (lane == (j & 31)) { if (j < 32) z0 = zj; else z1 = zj; }
        if (lane > j && lane < s) z0 = fma(-L[pidx(lane, j)], zj, z0);
        if (lane + 32 > j && lane + 32 < s) z1 = fma(-L[pidx(lane + 32, j)], zj, z1);
    }
    // backward: L^T x = z (row j of L is contiguous)
    for (int j = s - 1; j >= 0; --j) {
        const double src = j < 32 ? z0 : z1;
        const double xj = __shfl_sync(0xffffffffu, src, j & 31) * invd[j];
        if (lane == (j & 31)) { if (j < 32) z0 = xj; else z1 = xj; }
        const double* rowj = L + pidx(j, 0);
        if (lane < j) z0 = fma(-rowj[lane], xj, z0);
        if (lane + 32 < j) z1 = fma(-rowj[lane + 32], xj, z1);
    }
    if (lane < s) x[lane] = z0;
    if (lane + 32 < s) x[lane + 32] = z1;
}

// ------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------

enum { ST_IDLE = 0, ST_BASE, ST_PUMP, ST_PROBE, ST_WALK, ST_FINAL };

struct Slot {           // lives in the owner warp's registers (warp-uniform)
    int spec, ia, state, it, nq, ns;
    double alpha, mu, Q0, Q1, Q2, nuf, c0;
};

template <int NT, int T>
struct Smem {
    static constexpr int SP = NT * 8;
    static constexpr int STAGE = CK * NT * 64;          // doubles per staging buffer
    // offsets in doubles
    int pk;
    __host__ __device__ explicit Smem(int pk_) : pk(pk_) {}
    __host__ __device__ int stage(int i) const { return i * STAGE; }
    __host__ __device__ int J(int b) const { return 2 * STAGE + b * pk; }
    __host__ __device__ int L(int b) const { return 2 * STAGE + (T + b) * pk; }
    __host__ __device__ int vecs() const { return 2 * STAGE + 2 * T * pk; }
    // vectors: t[8][SP] then per-slot arrays [T][SP]
    __host__ __device__ int t() const { return vecs(); }
    __host__ __device__ int v(int b) const { return vecs() + 8 * SP + (0 * T + b) * SP; }
    __host__ __device__ int f(int b) const { return vecs() + 8 * SP + (1 * T + b) * SP; }
    __host__ __device__ int rhs(int b) const { return vecs() + 8 * SP + (2 * T + b) * SP; }
    __host__ __device__ int dvc(int b) const { return vecs() + 8 * SP + (3 * T + b) * SP; }
    __host__ __device__ int dvn(int b) const { return vecs() + 8 * SP + (4 * T + b) * SP; }
    __host__ __device__ int gt(int b) const { return vecs() + 8 * SP + (5 * T + b) * SP; }
    __host__ __device__ int y(int b) const { return vecs() + 8 * SP + (6 * T + b) * SP; }
    __host__ __device__ int u(int b) const { return vecs() + 8 * SP + (7 * T + b) * SP; }
    __host__ __device__ int invd(int b) const { return vecs() + 8 * SP + (8 * T + b) * SP; }
    __host__ __device__ int xi() const { return vecs() + 8 * SP + 9 * T * SP; }
    __host__ __device__ int lam() const { return xi() + SP; }
    __host__ __device__ int hch() const { return lam() + SP; }            // [FMAX][CK*8]
    __host__ __device__ int spart() const { return hch() + FMAX * CK * 8; }  // [NW][8]
    __host__ __device__ int res() const { return spart() + NW * 8; }      // chi2[8], S[8], Q[8], alpha[8]
    __host__ __device__ int ints() const { return res() + 32; }           // int region (64 ints)
    __host__ __device__ int total_doubles() const { return ints() + 32; }
};

// control block (ints) layout
enum { C_WANT = 0, C_WOUT = 8, C_SPEC = 16, C_SERVED = 24, C_HSLOT = 32, C_NH = 40, C_NACT = 41, C_ANYW = 42 };

// ------------------------------------------------------------------------------------------
// Z accumulation for one warp: row pair (IA, NT-1-IA), k-steps of parity e in one chunk
// ------------------------------------------------------------------------------------------
template <int NT, int IA>
__device__ __forceinline__ void zphase(const double* __restrict__ stage, const double* __restrict__ hch,
                                       int nh, int ntiles, int e, int offY, int lane,
                                       double (&zacc)[FMAX][NT + 1][2]) {
    constexpr int IB = NT - 1 - IA;
    static_assert(IB >= IA, "row pair order");
    for (int ktl = 0; ktl < ntiles; ++ktl) {
        const double* tile0 = stage + ktl * NT * 64 + offY;
        double fr[IB + 1];
#pragma unroll
        for (int jt = 0; jt <= IB; ++jt) fr[jt] = tile0[jt * 64];
#pragma unroll
        for (int f = 0; f < FMAX; ++f) {
            if (f < nh) {
                const double wl = hch[f * CK * 8 + ktl * 8 + 2 * (lane & 3) + e];
                double sc[IB + 1];
#pragma unroll
                for (int jt = 0; jt <= IB; ++jt) sc[jt] = wl * fr[jt];
#pragma unroll
                for (int jt = 0; jt <= IA; ++jt) dmma(zacc[f][jt], fr[IA], sc[jt]);
                if (IB != IA) {
#pragma unroll
                    for (int jt = 0; jt <= IB; ++jt) dmma(zacc[f][IA + 1 + jt], fr[IB], sc[jt]);
                }
            }
        }
    }
}

// store / add this warp's Z tiles into the packed buffer Zp (lower triangle, i,j < s)
template <int NT, int IA>
__device__ __forceinline__ void zstore(double* __restrict__ Zp, int s, int lane, bool add,
                                       const double (&zacc)[NT + 1][2]) {
    constexpr int IB = NT - 1 - IA;
    const int r = lane >> 2, c0 = 2 * (lane & 3);
#pragma unroll
    for (int pos = 0; pos <= NT; ++pos) {
        int I, Jt;
        if (pos <= IA) { I = IA; Jt = pos; }
        else { if (IB == IA) continue; I = IB; Jt = pos - IA - 1; }
        const int i = 8 * I + r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 8 * Jt + c0 + h;
            if (i < s && j <= i) {
                const int id = pidx(i, j);
                Zp[id] = add ? Zp[id] + zacc[pos][h] : zacc[pos][h];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int NT, int T>
__global__ void __launch_bounds__(NTHREADS, 1) sweep_kernel(const SweepArgs a) {
    constexpr int SP = NT * 8;
    constexpr int NPAIR = (NT + 1) / 2;
    static_assert(NPAIR <= NW / 2, "one row pair per warp");
    static_assert(T <= 8, "spectra are the M dimension of the MMA");
    extern __shared__ __align__(16) double sm[];
    const Smem<NT, T> L(a.pk);
    int* ctl = reinterpret_cast<int*>(sm + L.ints());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = a.n_sv;
    const double eps_nu = a.nu * 2.220446049250313e-16;

    // lane-constant fragment offsets inside a swizzled 64-element tile
    const int offX = tile_off(lane >> 2, 2 * (lane & 3));                 // 16-byte aligned pair
    const int offY0 = tile_off(2 * (lane & 3) + 0, lane >> 2);
    const int offY1 = tile_off(2 * (lane & 3) + 1, lane >> 2);

    for (int i = tid; i < SP; i += NTHREADS) {
        const double x = i < s ? a.xi[i] : 0.0;
        sm[L.xi() + i] = x;
        sm[L.lam() + i] = x * x;
    }
    for (int i = tid; i < 8 * SP; i += NTHREADS) sm[L.t() + i] = 0.0;
    if (tid < 64) ctl[tid] = 0;
    if (tid < 8) { ctl[C_WOUT + tid] = -1; ctl[C_SPEC + tid] = -1; }
    __syncthreads();

    Slot st;
    st.spec = -1; st.ia = 0; st.state = ST_IDLE; st.it = 0; st.nq = 0; st.ns = 0;
    st.alpha = 0; st.mu = 0; st.Q0 = 0; st.Q1 = 0; st.Q2 = 0; st.nuf = 0; st.c0 = 0;

    // ---- owner-warp helpers ------------------------------------------------------------
    auto fetch = [&](int b) {       // grab the next spectrum for slot b (owner warp only)
        int sp = 0;
        if (lane == 0) sp = atomicAdd(a.counter, 1);
        sp = __shfl_sync(0xffffffffu, sp, 0);
        if (sp >= a.B) {
            st.state = ST_IDLE; st.spec = -1;
            if (lane == 0) { ctl[C_WANT + b] = 0; ctl[C_WOUT + b] = -1; ctl[C_SPEC + b] = -1; }
            for (int i = lane; i < SP; i += 32) sm[L.t() + b * SP + i] = 0.0;
            return;
        }
        st.spec = sp; st.ia = 0; st.state = ST_BASE; st.it = 0; st.nq = 0; st.ns = 0;
        st.alpha = a.alpha[0]; st.mu = a.mu0; st.Q0 = nan(""); st.c0 = a.c0[sp];
        for (int i = lane; i < SP; i += 32) {
            const double v0 = i < s ? a.v0[i] : 0.0;
            sm[L.v(b) + i] = v0;
            sm[L.t() + b * SP + i] = v0;
            sm[L.gt(b) + i] = i < s ? a.gt[(size_t)sp * s + i] : 0.0;
        }
        if (lane == 0) { ctl[C_WANT + b] = 2; ctl[C_WOUT + b] = -1; ctl[C_SPEC + b] = sp; sm[L.res() + 24 + b] = st.alpha; }
    };
    // dest = (J + mu*shift)^-1 rhs ; false + dest = 0 on a failed factorisation
    auto solve = [&](int b, double mu, double* dest) -> bool {
        double* Lb = sm + L.L(b);
        const double* Jb = sm + L.J(b);
        const int np = pidx(s, 0);
        for (int i = lane; i < np; i += 32) Lb[i] = Jb[i];
        __syncwarp();
        for (int i = lane; i < s; i += 32) {
            const double sh = (a.variant == MX_VARIANT_BRYAN) ? mu / sm[L.lam() + i] : mu;
            Lb[pidx(i, i)] += sh;
        }
        __syncwarp();
        st.ns++;
        const bool ok = chol_packed(Lb, s, sm + L.invd(b), lane);
        if (!ok) { for (int i = lane; i < SP; i += 32) dest[i] = 0.0; __syncwarp(); return false; }
        chol_solve(Lb, sm + L.invd(b), s, sm + L.rhs(b), dest, lane);
        __syncwarp();
        return true;
    };
    auto set_trial = [&](int b, const double* dv, int want) {   // t = v - dv (dv may be null: t = v)
        for (int i = lane; i < SP; i += 32) {
            const double vv = sm[L.v(b) + i];
            sm[L.t() + b * SP + i] = (dv != nullptr && i < s) ? vv - dv[i] : vv;
        }
        if (lane == 0) ctl[C_WANT + b] = want;
    };
    auto finish_alpha = [&](int b, bool converged, int n_iter) {
        const size_t o = (size_t)st.spec * a.n_alpha + st.ia;
        double logp = nan("");
        if (a.want_prob) {
            // log p = -1/2 log det(I + eta Xi Z Xi / alpha) - Q - log alpha   (probabilities.py:76-85, Sylvester)
            double* M = sm + L.J(b);
            const double* Zp = sm + L.L(b);
            const int np = pidx(s, 0);
            for (int id = lane; id < np; id += 32) {
                // invert packed index
                int i = (int)((sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
                while (pidx(i + 1, 0) <= id) ++i;
                while (pidx(i, 0) > id) --i;
                const int j = id - pidx(i, 0);
                double m = a.eta * sm[L.xi() + i] * Zp[id] * sm[L.xi() + j] / st.alpha;
                if (i == j) m += 1.0;
                M[id] = m;
            }
            __syncwarp();
            const bool ok = chol_packed(M, s, sm + L.invd(b), lane);
            if (ok) {
                double ld = 0.0;
                for (int i = lane; i < s; i += 32) ld += log(M[pidx(i, i)]);
                ld = 2.0 * warp_sum(ld);
                logp = -0.5 * ld - st.Q1 - log(st.alpha);
            }
        }
        if (lane == 0) {
            a.o_chi2[o] = sm[L.res() + 0 + b];
            a.o_S[o] = sm[L.res() + 8 + b];
            a.o_Q[o] = st.Q1;
            if (a.o_logp) a.o_logp[o] = logp;
            if (a.o_niter) a.o_niter[o] = n_iter;
            if (a.o_nq) a.o_nq[o] = st.nq;
            if (a.o_ns) a.o_ns[o] = st.ns;
            if (a.o_status) a.o_status[o] = converged ? MX_STATUS_CONVERGED : 0;
        }
        if (a.o_v) for (int i = lane; i < s; i += 32) a.o_v[o * s + i] = sm[L.v(b) + i];
    };
    // Levenberg-Marquardt state machine: consume the result of the round just finished and set up
    // the next trial.  Mirrors levenberg_minimizer.py:143-248 statement by statement.
    auto advance = [&](int b) {
        if (st.state == ST_IDLE) return;
        if (!ctl[C_SERVED + b]) return;                // deferred Hessian round: ask again
        double Qt = sm[L.res() + 16 + b];
        double* dvc = sm + L.dvc(b);
        double* dvn = sm + L.dvn(b);
        int phase = st.state;
        if (phase == ST_FINAL) { fetch(b); return; }
        if (phase == ST_BASE) {
            st.Q1 = Qt; st.nq++;
            if (lane == 0) ctl[C_WOUT + b] = -1;
            double mf = 0.0;
            for (int i = lane; i < s; i += 32) mf = fmax(mf, fabs(sm[L.f(b) + i]));
            mf = warp_max(mf);
            // MaxDerivative(1e-4) | RelativeFunctionChange(1e-16)   (levenberg_minimizer.py:103-106)
            const bool conv = (mf < a.conv_maxd) || (fabs(fabs(st.Q0 - st.Q1) / st.Q1) < a.conv_relq);
            if ((conv && st.it >= a.miniter) || st.it >= a.maxiter) {
                const bool done_conv = st.it < a.maxiter ? conv : false;
                finish_alpha(b, done_conv, st.it < a.maxiter ? st.it + 1 : a.maxiter);
                const int prev = st.ia;
                st.ia++;
                if (st.ia < a.n_alpha) {
                    st.alpha = a.alpha[st.ia]; st.mu = a.mu0; st.Q0 = nan(""); st.it = 0; st.nq = 0; st.ns = 0;
                    st.state = ST_BASE;
                    set_trial(b, nullptr, 2);
                    if (lane == 0) sm[L.res() + 24 + b] = st.alpha;
                } else {
                    st.state = ST_FINAL;
                    set_trial(b, nullptr, 1);
                }
                if (lane == 0) ctl[C_WOUT + b] = (a.o_A != nullptr) ? prev : -1;
                return;
            }
            st.Q0 = st.Q1;
            st.state = ST_PUMP;
            if (solve(b, st.mu, dvc)) { set_trial(b, dvc, 1); return; }
            st.Q1 = nan("");
            phase = -ST_PUMP;      // fall into the pump loop without consuming a result
        }
        if (phase == ST_PUMP) { st.Q1 = Qt; st.nq++; }
        if (phase == ST_PUMP || phase == -ST_PUMP) {
            // while (Q1 > Q0 or isnan(Q1)) and mu < max_mu      (levenberg_minimizer.py:203-206)
            while ((st.Q1 > st.Q0 || isnan(st.Q1)) && st.mu < a.max_mu) {
                st.mu *= a.nu;
                if (solve(b, st.mu, dvc)) { st.state = ST_PUMP; set_trial(b, dvc, 1); return; }
                st.Q1 = nan("");
            }
            st.state = ST_PROBE;                         // dv2 = solve(J + nu*mu)   (:209-210)
            if (solve(b, a.nu * st.mu, dvn)) { set_trial(b, dvn, 1); return; }
            Qt = nan("");
            phase = -ST_PROBE;
        }
        if (phase == ST_PROBE) st.nq++;
        if (phase == ST_PROBE || phase == -ST_PROBE) {
            st.Q2 = Qt;
            if (st.Q2 < st.Q1) {                         // (:214-218)
                st.nuf = a.nu; st.mu *= a.nu; st.Q2 = st.Q1;
            } else {                                     // (:221-224)
                st.nuf = 1.0 / a.nu; st.mu /= st.nuf;
                for (int i = lane; i < SP; i += 32) dvn[i] = dvc[i];
                __syncwarp();
            }
            st.Q1 = INFINITY;
            phase = -ST_WALK;
        }
        if (phase == ST_WALK) { st.Q2 = Qt; st.nq++; }
        if (phase == ST_WALK || phase == -ST_WALK) {
            // while Q2 < Q1 and mu < max_mu and mu > nu*eps      (:227-233)
            while (st.Q2 < st.Q1 && st.mu < a.max_mu && st.mu > eps_nu) {
                st.Q1 = st.Q2;
                for (int i = lane; i < SP; i += 32) dvc[i] = dvn[i];
                __syncwarp();
                st.mu *= st.nuf;
                if (solve(b, st.mu, dvn)) { st.state = ST_WALK; set_trial(b, dvn, 1); return; }
                st.Q2 = nan("");
            }
            // v -= dv ; func_val = function(v)                   (:239-243)
            for (int i = lane; i < s; i += 32) sm[L.v(b) + i] -= dvc[i];
            __syncwarp();
            st.it++;
            st.state = ST_BASE;
            set_trial(b, nullptr, 2);
        }
    };

    if (warp < T) fetch(warp);

    const int nch = (a.n_kt + CK - 1) / CK;
    const bool pm = a.variant == MX_VARIANT_PLUSMINUS;
    int round = 0;

    while (true) {
        __syncthreads();
        // ---- plan the round ----------------------------------------------------------------
        if (tid == 0) {
            int nh = 0, nact = 0, anyw = 0;
            for (int q = 0; q < T; ++q) {
                const int b = (q + round) % T;          // rotate priority so nobody starves
                const int w = ctl[C_WANT + b];
                int served = 0;
                if (w == 1) served = 1;
                else if (w == 2) { if (nh < FMAX) { ctl[C_HSLOT + nh] = b; ++nh; served = 1; } }
                ctl[C_SERVED + b] = served;
                if (w) ++nact;
                if (served && ctl[C_WOUT + b] >= 0) anyw = 1;
            }
            ctl[C_NH] = nh; ctl[C_NACT] = nact; ctl[C_ANYW] = anyw;
        }
        __syncthreads();
        if (ctl[C_NACT] == 0) break;
        const int nh = ctl[C_NH];
        ++round;

        // ---- the pass over V' -----------------------------------------------------------------
        double tA[NT][2];
        {
            const int b = lane >> 2;
            const bool live = b < T && ctl[C_SERVED + b];
#pragma unroll
            for (int jt = 0; jt < NT; ++jt) {
                const double2 tv = *reinterpret_cast<const double2*>(sm + L.t() + b * SP + 8 * jt + 2 * (lane & 3));
                tA[jt][0] = live ? tv.x : 0.0;
                tA[jt][1] = live ? tv.y : 0.0;
            }
        }
        double yacc[NT][2];
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) { yacc[jt][0] = 0.0; yacc[jt][1] = 0.0; }
        double zacc[FMAX][NT + 1][2];
#pragma unroll
        for (int f = 0; f < FMAX; ++f)
#pragma unroll
            for (int p = 0; p <= NT; ++p) { zacc[f][p][0] = 0.0; zacc[f][p][1] = 0.0; }
        double sacc = 0.0;
        const int myb = lane >> 2;
        const int wout = (myb < T && ctl[C_SERVED + myb]) ? ctl[C_WOUT + myb] : -1;
        const int myspec = myb < T ? ctl[C_SPEC + myb] : -1;
        int hsel[FMAX];
#pragma unroll
        for (int f = 0; f < FMAX; ++f) hsel[f] = (f < nh) ? ctl[C_HSLOT + f] : -1;

        auto issue = [&](int c) {
            const int t0 = c * CK;
            const int nt = min(CK, a.n_kt - t0);
            const double* src = a.Vt + (size_t)t0 * NT * 64;
            double* dst = sm + L.stage(c & 1);
            const int n16 = nt * NT * 32;               // 16-byte packets
            for (int i = tid; i < n16; i += NTHREADS) cp_async16(dst + 2 * i, src + 2 * i);
            cp_async_commit();
        };
        issue(0);
        for (int c = 0; c < nch; ++c) {
            cp_async_wait_all();
            __syncthreads();
            if (c + 1 < nch) issue(c + 1);
            const double* stage = sm + L.stage(c & 1);
            const int kt = c * CK + warp;
            const int ntiles = min(CK, a.n_kt - c * CK);
            if (kt < a.n_kt) {
                const double* tile0 = stage + warp * NT * 64;
                double C[2] = {0.0, 0.0};
#pragma unroll
                for (int jt = 0; jt < NT; ++jt) {
                    const double2 vv = *reinterpret_cast<const double2*>(tile0 + jt * 64 + offX);
                    dmma(C, tA[jt][0], vv.x);
                    dmma(C, tA[jt][1], vv.y);
                }
                double Hv[2], Wv[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int k = kt * 8 + 2 * (lane & 3) + i;
                    const double Dk = k < a.n_omega ? __ldg(a.D + k) : 0.0;
                    const double x = C[i];
                    const double ex = exp(x);
                    double H, W, st_;
                    if (!pm) {
                        // H = D e^x ; S += H - D - H log(H/D), safelog clamp at 1e-100  (functions.py:53-56,508-510)
                        H = Dk * ex; W = H;
                        const double lg = (ex <= 1e-100) ? -230.25850929940458 : x;
                        st_ = H - Dk - H * lg;
                    } else {
                        // H = D (e^x - e^-x) ; w = D (e^x + e^-x) ; S = S_n(H+) + S_n(H-)   (functions.py:544-564,778-786)
                        const double em = exp(-x);
                        const double Hp = Dk * ex, Hm = Dk * em;
                        H = Hp - Hm; W = Hp + Hm;
                        const double lp = (ex <= 1e-100) ? -230.25850929940458 : x;
                        const double lm = (em <= 1e-100) ? -230.25850929940458 : -x;
                        st_ = (Hp - Dk - Hp * lp) + (Hm - Dk - Hm * lm);
                    }
                    if (Dk == 0.0) { H = 0.0; W = 0.0; st_ = 0.0; }   // padded rows
                    sacc += st_;
                    Hv[i] = H; Wv[i] = W;
                    if (wout >= 0 && k < a.n_omega)
                        a.o_A[((size_t)myspec * a.n_alpha + wout) * a.n_omega + k] = H / __ldg(a.delta + k);   // functions.py:947-952
                }
#pragma unroll
                for (int f = 0; f < FMAX; ++f) {
                    if (myb == hsel[f]) {
                        sm[L.hch() + f * CK * 8 + warp * 8 + 2 * (lane & 3) + 0] = Wv[0];
                        sm[L.hch() + f * CK * 8 + warp * 8 + 2 * (lane & 3) + 1] = Wv[1];
                    }
                }
#pragma unroll
                for (int jt = 0; jt < NT; ++jt) {
                    dmma(yacc[jt], Hv[0], tile0[jt * 64 + offY0]);
                    dmma(yacc[jt], Hv[1], tile0[jt * 64 + offY1]);
                }
            }
            if (nh > 0) {
                __syncthreads();
                const int e = warp >> 2;                 // k-step parity handled by this warp
                const int offY = e ? offY1 : offY0;
                const double* hch = sm + L.hch();
                switch (warp & 3) {
                    case 0: zphase<NT, 0>(stage, hch, nh, ntiles, e, offY, lane, zacc); break;
                    case 1: if (NPAIR > 1) zphase<NT, (NPAIR > 1 ? 1 : 0)>(stage, hch, nh, ntiles, e, offY, lane, zacc); break;
                    case 2: if (NPAIR > 2) zphase<NT, (NPAIR > 2 ? 2 : 0)>(stage, hch, nh, ntiles, e, offY, lane, zacc); break;
                    case 3: if (NPAIR > 3) zphase<NT, (NPAIR > 3 ? 3 : 0)>(stage, hch, nh, ntiles, e, offY, lane, zacc); break;
                }
            }
        }
        __syncthreads();                                 // staging buffers are free: reuse as ypart

        // ---- reductions ---------------------------------------------------------------------
        double* ypart = sm + L.stage(0);                 // [NW][8][SP]
        {
            const int b = lane >> 2;
#pragma unroll
            for (int jt = 0; jt < NT; ++jt) {
                double2 o; o.x = yacc[jt][0]; o.y = yacc[jt][1];
                *reinterpret_cast<double2*>(ypart + (warp * 8 + b) * SP + 8 * jt + 2 * (lane & 3)) = o;
            }
            double sq = sacc;
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            if ((lane & 3) == 0) sm[L.spart() + warp * 8 + b] = sq;
        }
        auto zstore_all = [&](bool add) {
#pragma unroll
            for (int f = 0; f < FMAX; ++f) {
                if (f < nh) {
                    double* Zp = sm + L.L(hsel[f]);
                    switch (warp & 3) {
                        case 0: zstore<NT, 0>(Zp, s, lane, add, zacc[f]); break;
                        case 1: if (NPAIR > 1) zstore<NT, (NPAIR > 1 ? 1 : 0)>(Zp, s, lane, add, zacc[f]); break;
                        case 2: if (NPAIR > 2) zstore<NT, (NPAIR > 2 ? 2 : 0)>(Zp, s, lane, add, zacc[f]); break;
                        case 3: if (NPAIR > 3) zstore<NT, (NPAIR > 3 ? 3 : 0)>(Zp, s, lane, add, zacc[f]); break;
                    }
                }
            }
        };
        if (nh > 0 && warp < 4) zstore_all(false);
        __syncthreads();
        for (int i = tid; i < T * SP; i += NTHREADS) {
            const int b = i / SP, j = i - b * SP;
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) acc += ypart[(w * 8 + b) * SP + j];
            sm[L.y(b) + j] = acc;
        }
        if (nh > 0 && warp >= 4) zstore_all(true);
        __syncthreads();

        // ---- chi2, S, Q (owner warps) ------------------------------------------------------------
        if (warp < T && ctl[C_SERVED + warp]) {
            const int b = warp;
            double c2 = 0.0;
            for (int i = lane; i < s; i += 32) {
                const double rr = sm[L.xi() + i] * sm[L.y(b) + i] - sm[L.gt(b) + i];
                c2 = fma(rr, rr, c2);
                // u = g + alpha v,  g = eta Xi rr      (maxent_cost_function.py:85-93 projected)
                sm[L.u(b) + i] = a.eta * sm[L.xi() + i] * rr + st.alpha * sm[L.v(b) + i];
            }
            c2 = warp_sum(c2) + st.c0;
            double S = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) S += sm[L.spart() + w * 8 + b];
            const double Q = 0.5 * c2 * a.eta - st.alpha * S;   // maxent_cost_function.py:82
            if (lane == 0) { sm[L.res() + 0 + b] = c2; sm[L.res() + 8 + b] = S; sm[L.res() + 16 + b] = Q; }
        }
        __syncthreads();

        // ---- gradient and Hessian for the served Hessian slots (all threads) ---------------------
        for (int f = 0; f < nh; ++f) {
            const int b = ctl[C_HSLOT + f];
            const double* Zp = sm + L.L(b);
            const double* u = sm + L.u(b);
            const double alpha = sm[L.res() + 24 + b];
            if (a.variant == MX_VARIANT_BRYAN) {
                // f = g + alpha v ; (eta Lambda Z + mu) dv = f  <=>  (eta Z + mu/Lambda) dv = f/Lambda
                for (int i = tid; i < s; i += NTHREADS) {
                    sm[L.f(b) + i] = u[i];
                    sm[L.rhs(b) + i] = u[i] / sm[L.lam() + i];
                }
                const int np = pidx(s, 0);
                for (int i = tid; i < np; i += NTHREADS) sm[L.J(b) + i] = a.eta * Zp[i];
            } else {
                for (int i = warp; i < s; i += NW) {            // f = Z u
                    double acc = 0.0;
                    for (int k = lane; k < s; k += 32) acc = fma(Zp[sidx(i, k)], u[k], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) { sm[L.f(b) + i] = acc; sm[L.rhs(b) + i] = acc; }
                }
                // J = eta Z Lambda Z + alpha Z   (maxent_cost_function.py:161-162 in singular space)
                for (int i = warp; i < s; i += NW) {
                    for (int j = lane; j <= i; j += 32) {
                        double acc = 0.0;
                        for (int k = 0; k < s; ++k) acc = fma(Zp[sidx(k, i)] * sm[L.lam() + k], Zp[sidx(k, j)], acc);
                        sm[L.J(b) + pidx(i, j)] = a.eta * acc + alpha * Zp[pidx(i, j)];
                    }
                }
            }
        }
        __syncthreads();

        // ---- Levenberg-Marquardt bookkeeping and the next solve (owner warps) ----------------------
        if (warp < T) advance(warp);
    }
}

// ------------------------------------------------------------------------------------------
// host-side launch
// ------------------------------------------------------------------------------------------
template <int NT, int T>
int launch_sweep(const SweepArgs& a, cudaStream_t stream, bool query, int* o_t, int* o_smem) {
    const Smem<NT, T> L(a.pk);
    const size_t bytes = (size_t)L.total_doubles() * sizeof(double);
    if (o_t) *o_t = T;
    if (o_smem) *o_smem = (int)bytes;
    if (query) return MX_OK;
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<NT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return MX_ERR_CUDA;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = (a.B + T - 1) / T;
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    sweep_kernel<NT, T><<<grid, NTHREADS, bytes, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? MX_OK : MX_ERR_CUDA;
}

template <int NT>
int pick_T(const SweepArgs& a, cudaStream_t stream, bool query, int* o_t, int* o_smem) {
    constexpr size_t LIMIT = 227 * 1024;
#define MX_TRY(TT)                                                                         \
    if ((size_t)Smem<NT, TT>(a.pk).total_doubles() * sizeof(double) <= LIMIT)              \
        return launch_sweep<NT, TT>(a, stream, query, o_t, o_smem);
    MX_TRY(8) MX_TRY(7) MX_TRY(6) MX_TRY(5) MX_TRY(4) MX_TRY(3) MX_TRY(2) MX_TRY(1)
#undef MX_TRY
    return MX_ERR_UNSUPPORTED;
}

}  // namespace mx
