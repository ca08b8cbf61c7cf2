// fam_sphere.cuh -- NonnegPCA on Sphere(n): min -x'Zx, s_i = x_i + eps > 0
// (src/NonnegPCA/coordinator.py:37-95; closed forms SURVEY.md App. A.1).
//
// One warp owns one instance.  S = Z + Z' (n x n, symmetric) sits in shared memory for the
// whole solve; vectors are WVec<K> in the pair layout described below, n <= 32*K.
//
//   grad f      = P_x(-Sx),   P_x u = u - <x,u> x        (pymanopt Sphere.projection)
//   Hess L[v]   = P_x(-Sv) + (x'Sx) v + (y'x) v           (RIPTRM.py:491-523 + Sphere.ehess2rhess)
//   G_x(w)      = P_x(w) ;  G*_x[v]_i = v_i - x_i <x,v>   (:525-571)
//   Hw[v]       = Hess L[v] + G_x((y/s) * G*_x[v])        (:729)
//   c           = grad f - G_x(mu / s)                     (:730)
#pragma once
#include "solver_warp.cuh"
#include "tmem.cuh"

namespace riptrm {

// Vector layout ("pair layout"): lane l owns the adjacent elements e = 64*(k>>1) + 2*l + (k&1), k = 0..K-1,
// so one LDS.128 per lane fetches a lane's two entries of a row of S (shared-memory instruction issue, one
// warp-wide load per 4 cycles per scheduler, is what bounds a single warp's S.v: 2x fewer instructions than
// LDS.64 with the lane-strided layout; measured with scripts/microbench.cu).  K = 2 for n <= 64, 4 for n <= 128.
// NFIX > 0 fixes n at compile time (the reference's dim = 50: S.v fully unrolled, no remainder loops).
// TM = true keeps S in Tensor Memory instead of shared memory (n = 50 only): lane l's two columns of S are 200
// consecutive 32-bit TMEM columns of the warp's own TMEM lane l, fetched with tcgen05.ld.32x32b (SASS LDTM).  With
// every SM full, S.v from shared memory costs ~1890 cycles (the shared-memory pipe is the bottleneck: 83 % busy),
// from TMEM ~630 (scripts/tmem_test.cu); the values and their order are the same, so results stay bit-identical.
// RS = true takes the two merged reductions of a tCG iteration through shared memory (wsum_smem: same trees, fewer
// instructions); Ctx::ws then holds the addresses of kWarpScratchDoubles doubles of this warp
// (the broadcast operand of S.v moves there too).
template <int K_, int NFIX = 0, bool TM = false, int TCH = 64, bool RS = false>
struct SphereFam {
    static_assert(!TM || (NFIX == 50 && K_ == 2), "the TMEM layout is built for n = 50");
    static constexpr int K = K_;
    static constexpr int MK = K_;
    static constexpr bool kTcgReturnsHw = true;   // tcg() returns Hw[eta] accumulated beside eta (solver_warp.cuh inner_step)
    static constexpr bool kMergedStepDots = true; // step_dots / gadj_given / retract_given below
    static_assert(K_ == 2 || K_ == 4, "pair layout: K is 2 or 4");
    using Vec = WVec<K>;
    using CVec = WVec<K>;

    struct Ctx {
        const double* S;  // shared memory, row-major n x ns (ns = n rounded up to even, + tail padding), symmetric
        double* vbuf;     // shared memory, 32*K doubles, staging for the broadcast operand
        int n, ns;
        double eps;
        bool embedded;    // 'is_euclidean_embedded'
        uint32_t taddr;   // TM: TMEM address of this warp's copy of S (lane field = 32 * (warp % 4))
        WarpScratch ws;   // RS: shared-space addresses of the warp scratch (kWarpScratchDoubles doubles, 16-byte aligned)
    };
    struct Pt {
        Vec x;
        CVec s;
        double cost;
        Vec Sx;
        double xSx;
    };

    // TM: move S from the staging buffer (shared memory, row-major) into this warp's TMEM lanes:
    // TMEM column 4j + 2k (+1 for the high word) of lane l holds S[j][2l + k]
    static __device__ __forceinline__ void stage_to_tmem(const Ctx& c) {
        const double2* row = reinterpret_cast<const double2*>(c.S) + lane_id();
        constexpr int H = NFIX > 0 ? NFIX / 2 : 1;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            uint32_t r[64];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const double2 sv = row[(16 * ch + jj) * H];
                r[4 * jj + 0] = (uint32_t)__double2loint(sv.x);
                r[4 * jj + 1] = (uint32_t)__double2hiint(sv.x);
                r[4 * jj + 2] = (uint32_t)__double2loint(sv.y);
                r[4 * jj + 3] = (uint32_t)__double2hiint(sv.y);
            }
            tmem::tmem_st_x64(c.taddr + 64 * ch, r);
        }
        uint32_t r[8];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const double2 sv = row[(48 + jj) * H];
            r[4 * jj + 0] = (uint32_t)__double2loint(sv.x);
            r[4 * jj + 1] = (uint32_t)__double2hiint(sv.x);
            r[4 * jj + 2] = (uint32_t)__double2loint(sv.y);
            r[4 * jj + 3] = (uint32_t)__double2hiint(sv.y);
        }
        tmem::tmem_st_x8(c.taddr + 192, r);
        tmem::wait_st();
    }

    static __device__ __forceinline__ Vec matvec_tmem(const Ctx& c, const Vec& v) {
        const int lane = lane_id();
        if (TCH == 16) sts_v2f64_off<0>(c.ws.wl, v.v[0], v.v[1]);
        else reinterpret_cast<double2*>(c.vbuf)[lane] = make_double2(v.v[0], v.v[1]);
        __syncwarp();
        double a0x = 0.0, a0y = 0.0, a1x = 0.0, a1y = 0.0;
        if (TCH == 16) {
            // 16-column chunks (the 128-register kernel: a chunk's S registers are half as many again; same order of
            // additions), operand broadcast through the warp scratch with explicit shared-space addresses
            static_for<0, 12>([&](auto chc) {
                constexpr int ch = decltype(chc)::value;
                uint32_t r[16];
                tmem::tmem_ld_x16(c.taddr + 16 * ch, r);
                tmem::wait_ld();
                const double2 v0 = lds_v2f64_off<32 * ch>(c.ws.wb), v1 = lds_v2f64_off<32 * ch + 16>(c.ws.wb);
                a0x = fma(__hiloint2double((int)r[1], (int)r[0]), v0.x, a0x);
                a0y = fma(__hiloint2double((int)r[3], (int)r[2]), v0.x, a0y);
                a1x = fma(__hiloint2double((int)r[5], (int)r[4]), v0.y, a1x);
                a1y = fma(__hiloint2double((int)r[7], (int)r[6]), v0.y, a1y);
                a0x = fma(__hiloint2double((int)r[9], (int)r[8]), v1.x, a0x);
                a0y = fma(__hiloint2double((int)r[11], (int)r[10]), v1.x, a0y);
                a1x = fma(__hiloint2double((int)r[13], (int)r[12]), v1.y, a1x);
                a1y = fma(__hiloint2double((int)r[15], (int)r[14]), v1.y, a1y);
            });
            uint32_t r[8];
            tmem::tmem_ld_x8(c.taddr + 192, r);
            tmem::wait_ld();
            const double2 vj = lds_v2f64_off<8 * 48>(c.ws.wb);
            a0x = fma(__hiloint2double((int)r[1], (int)r[0]), vj.x, a0x);
            a0y = fma(__hiloint2double((int)r[3], (int)r[2]), vj.x, a0y);
            a1x = fma(__hiloint2double((int)r[5], (int)r[4]), vj.y, a1x);
            a1y = fma(__hiloint2double((int)r[7], (int)r[6]), vj.y, a1y);
            __syncwarp();
            Vec out;
            out.v[0] = active(c, 0) ? (a0x + a1x) : 0.0;
            out.v[1] = active(c, 1) ? (a0y + a1y) : 0.0;
            return out;
        } else if (TCH == 32) {
            // 32-column chunks for the 128-register kernel (same order of additions)
#pragma unroll
            for (int ch = 0; ch < 6; ++ch) {
                uint32_t r[32];
                tmem::tmem_ld_x32(c.taddr + 32 * ch, r);
                tmem::wait_ld();
#pragma unroll
                for (int jj = 0; jj < 8; jj += 2) {
                    const double2 vj = *reinterpret_cast<const double2*>(c.vbuf + 8 * ch + jj);
                    a0x = fma(__hiloint2double((int)r[4 * jj + 1], (int)r[4 * jj + 0]), vj.x, a0x);
                    a0y = fma(__hiloint2double((int)r[4 * jj + 3], (int)r[4 * jj + 2]), vj.x, a0y);
                    a1x = fma(__hiloint2double((int)r[4 * jj + 5], (int)r[4 * jj + 4]), vj.y, a1x);
                    a1y = fma(__hiloint2double((int)r[4 * jj + 7], (int)r[4 * jj + 6]), vj.y, a1y);
                }
            }
        } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            uint32_t r[64];
            tmem::tmem_ld_x64(c.taddr + 64 * ch, r);
            tmem::wait_ld();
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) {
                const double2 vj = *reinterpret_cast<const double2*>(c.vbuf + 16 * ch + jj);
                a0x = fma(__hiloint2double((int)r[4 * jj + 1], (int)r[4 * jj + 0]), vj.x, a0x);
                a0y = fma(__hiloint2double((int)r[4 * jj + 3], (int)r[4 * jj + 2]), vj.x, a0y);
                a1x = fma(__hiloint2double((int)r[4 * jj + 5], (int)r[4 * jj + 4]), vj.y, a1x);
                a1y = fma(__hiloint2double((int)r[4 * jj + 7], (int)r[4 * jj + 6]), vj.y, a1y);
            }
        }
        }
        {
            uint32_t r[8];
            tmem::tmem_ld_x8(c.taddr + 192, r);
            tmem::wait_ld();
            const double2 vj = *reinterpret_cast<const double2*>(c.vbuf + 48);
            a0x = fma(__hiloint2double((int)r[1], (int)r[0]), vj.x, a0x);
            a0y = fma(__hiloint2double((int)r[3], (int)r[2]), vj.x, a0y);
            a1x = fma(__hiloint2double((int)r[5], (int)r[4]), vj.y, a1x);
            a1y = fma(__hiloint2double((int)r[7], (int)r[6]), vj.y, a1y);
        }
        __syncwarp();
        Vec out;
        out.v[0] = active(c, 0) ? (a0x + a1x) : 0.0;
        out.v[1] = active(c, 1) ? (a0y + a1y) : 0.0;
        return out;
    }
    struct Step {
        Vec c;
        CVec ys;       // y / s
        double kappa;  // x'Sx + y'x
    };

    static __device__ __forceinline__ int dimn(const Ctx& c) { return NFIX > 0 ? NFIX : c.n; }
    static __device__ __forceinline__ int stride(const Ctx& c) { return NFIX > 0 ? ((NFIX + 1) & ~1) : c.ns; }
    static __device__ __forceinline__ int elem(int k) { return 64 * (k >> 1) + 2 * lane_id() + (k & 1); }
    static __device__ __forceinline__ bool active(const Ctx& c, int k) { return elem(k) < dimn(c); }
    static __device__ __forceinline__ bool cactive(const Ctx& c, int k) { return active(c, k); }
    static __device__ __forceinline__ int dim(const Ctx& c) { return dimn(c) - 1; }
    static __device__ __forceinline__ int num_constraints(const Ctx& c) { return dimn(c); }
    static __device__ __forceinline__ double typical_dist(const Ctx&) { return 3.141592653589793; }
    static __device__ __forceinline__ bool domain_ok(const Ctx&, const Pt&) { return true; }

    // out = S v.  S symmetric, so (S v)_e = sum_j S[j][e] v_j: lane-contiguous (conflict-free) 16-byte reads of
    // row j, v_j broadcast from shared memory.  Two accumulators per element (even j / odd j) halve the
    // dependent-FMA chain; their sum order is part of the arithmetic specification.
    static __device__ __forceinline__ Vec matvec(const Ctx& c, const Vec& v) {
        if (TM) return matvec_tmem(c, v);
        const int lane = lane_id();
        const int n = dimn(c);
        double2* vb2 = reinterpret_cast<double2*>(c.vbuf);
#pragma unroll
        for (int q = 0; q < K / 2; ++q) vb2[32 * q + lane] = make_double2(v.v[2 * q], v.v[2 * q + 1]);
        __syncwarp();
        double a0[K], a1[K];
#pragma unroll
        for (int k = 0; k < K; ++k) a0[k] = a1[k] = 0.0;
        const int ns2 = stride(c) >> 1;
        const double2* row = reinterpret_cast<const double2*>(c.S) + lane;
        int j = 0;
#pragma unroll(NFIX > 0 ? 32 : 5)
        for (; j + 1 < n; j += 2) {
            const double2 vj = *reinterpret_cast<const double2*>(c.vbuf + j);
#pragma unroll
            for (int q = 0; q < K / 2; ++q) {
                const double2 s0 = row[32 * q], s1 = row[ns2 + 32 * q];
                a0[2 * q] = fma(s0.x, vj.x, a0[2 * q]);
                a0[2 * q + 1] = fma(s0.y, vj.x, a0[2 * q + 1]);
                a1[2 * q] = fma(s1.x, vj.y, a1[2 * q]);
                a1[2 * q + 1] = fma(s1.y, vj.y, a1[2 * q + 1]);
            }
            row += 2 * ns2;
        }
        if (j < n) {
            const double vj = c.vbuf[j];
#pragma unroll
            for (int q = 0; q < K / 2; ++q) {
                const double2 s0 = row[32 * q];
                a0[2 * q] = fma(s0.x, vj, a0[2 * q]);
                a0[2 * q + 1] = fma(s0.y, vj, a0[2 * q + 1]);
            }
        }
        __syncwarp();
        Vec out;
#pragma unroll
        for (int k = 0; k < K; ++k) out.v[k] = active(c, k) ? (a0[k] + a1[k]) : 0.0;
        return out;
    }

    static __device__ __forceinline__ void eval_point(const Ctx& c, const Vec& x, Pt& pt) {
        pt.x = x;
        pt.Sx = matvec(c, x);
        pt.xSx = wdot(x, pt.Sx);
        pt.cost = -0.5 * pt.xSx;  // -x'Zx = -x'Sx/2
#pragma unroll
        for (int k = 0; k < K; ++k) pt.s.v[k] = active(c, k) ? (x.v[k] + c.eps) : 0.0;
    }

    // per-lane partial of <a, b>_x (the Sphere metric is the Euclidean one); wsum() of it is the inner product
    static __device__ __forceinline__ double inner_partial(const Ctx&, const Pt&, const Vec& a, const Vec& b) {
        return wdot_partial(a, b);
    }
    static __device__ __forceinline__ double inner(const Ctx&, const Pt&, const Vec& a, const Vec& b) {
        return wdot(a, b);
    }

    static __device__ __forceinline__ Vec project(const Ctx&, const Pt& pt, const Vec& v) {
        const double a = wdot(pt.x, v);
        Vec r;
#pragma unroll
        for (int k = 0; k < K; ++k) r.v[k] = v.v[k] - a * pt.x.v[k];
        return r;
    }

    static __device__ __forceinline__ void begin_step(const Ctx& c, const Pt& pt, const CVec& y, double mu, Step& st) {
        Vec w;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool on = active(c, k);
            w.v[k] = on ? mu * (1.0 / pt.s.v[k]) : 0.0;
            st.ys.v[k] = on ? y.v[k] / pt.s.v[k] : 0.0;
        }
        double xw = wdot_partial(pt.x, w), yx = wdot_partial(y, pt.x);
        wsum2(xw, yx);
        st.kappa = pt.xSx + yx;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double gradf = -pt.Sx.v[k] + pt.xSx * pt.x.v[k];  // P_x(-Sx); <x,-Sx> = -x'Sx exactly
            const double Gw = w.v[k] - xw * pt.x.v[k];             // G_x(mu/s) = P_x(mu/s)
            st.c.v[k] = gradf - Gw;
        }
    }

    static __device__ __forceinline__ CVec gadj(const Ctx& c, const Pt& pt, const Vec& v) {
        if (c.embedded) return v;
        const double b = wdot(pt.x, v);
        CVec g;
#pragma unroll
        for (int k = 0; k < K; ++k) g.v[k] = v.v[k] - pt.x.v[k] * b;
        return g;
    }

    // The five inner products of a trust-region iteration that depend on dx alone, in one reduction round (inner_step):
    // <dx,dx>, <x,dx>, <x+dx,x+dx>, <Hdx,dx>, <c,dx>; per value the tree of wdot(), hence the bits of the separate calls.
    static __device__ __forceinline__ void step_dots(const Ctx&, const Pt& pt, const Step& st, const Vec& dx, const Vec& Hdx,
                                                     double (&d)[5]) {
        Vec a;
#pragma unroll
        for (int k = 0; k < K; ++k) a.v[k] = pt.x.v[k] + dx.v[k];
        d[0] = wdot_partial(dx, dx);
        d[1] = wdot_partial(pt.x, dx);
        d[2] = wdot_partial(a, a);
        d[3] = wdot_partial(Hdx, dx);
        d[4] = wdot_partial(st.c, dx);
        wsumN<5>(d);
    }
    static __device__ __forceinline__ CVec gadj_given(const Ctx& c, const Pt& pt, const Vec& v, double b) {
        if (c.embedded) return v;
        CVec g;
#pragma unroll
        for (int k = 0; k < K; ++k) g.v[k] = v.v[k] - pt.x.v[k] * b;
        return g;
    }
    static __device__ __forceinline__ Vec retract_given(const Ctx&, const Pt& pt, const Vec& dx, double aa) {
        Vec a;
        const double nrm = sqrt(aa);
#pragma unroll
        for (int k = 0; k < K; ++k) a.v[k] = (pt.x.v[k] + dx.v[k]) / nrm;
        return a;
    }

    static __device__ __forceinline__ Vec Hw(const Ctx& c, const Pt& pt, const CVec&, const Step& st, const Vec& v) {
        const Vec Sv = matvec(c, v);
        double a = wdot_partial(pt.x, Sv), b = wdot_partial(pt.x, v);
        wsum2(a, b);
        Vec t, out;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double ga = c.embedded ? v.v[k] : (v.v[k] - pt.x.v[k] * b);
            t.v[k] = st.ys.v[k] * ga;
        }
        const double d = wdot(pt.x, t);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double hl = (-Sv.v[k] + a * pt.x.v[k]) + st.kappa * v.v[k];
            const double g = t.v[k] - d * pt.x.v[k];
            out.v[k] = hl + g;
        }
        return out;
    }

    // Steihaug-Toint tCG (RIPTRM.py:41-216) with merged reductions.  A warp-wide fp64 sum costs ~185 cycles
    // of pure latency on B200 whether it carries one value or six (scripts/microbench.cu), and the loop in the
    // reference's operation order needs five dependent ones per iteration.  Here <x,t>, <delta, Hw delta> and the
    // coefficient of the re-projection of delta (:210) are assembled from inner products that do not depend on
    // each other: with w = x*(y/s), q = <w,x>,
    //     d      = <x, (y/s)*(delta - b x)>            = <w,delta> - b q
    //     d_Hd   = <delta, P(-S delta) + kappa delta + P((y/s)*G*[delta])>
    //            = -<delta,S delta> + a b + kappa <delta,delta> + (<delta,(y/s)*delta> - b <w,delta>) - d b
    //     <x, -r' + beta delta> = -<x,r'> + beta b
    // so an iteration is S.delta, one 6-value reduction, the step, one 4-value reduction.  Same algorithm, same
    // exits; rounding differs from the reference order by the usual few ulp (oracle/c/riptrm_det.c implements
    // exactly this arithmetic and tests/test_oracle_c.py compares both orders and the reference's golden run).
    static __device__ __forceinline__ TcgResult tcg(const Ctx& ctx, const DevOpts& o, const Pt& pt, const CVec& y,
                                                    const Step& st, double Delta, Vec& eta, Vec& Heta) {
        (void)y;
        eta = wzero<K>();
        Heta = wzero<K>();                                     // :47
        Vec r = st.c, delta, wv;                               // :48
#pragma unroll
        for (int k = 0; k < K; ++k) {
            delta.v[k] = -r.v[k];                              // :71
            wv.v[k] = pt.x.v[k] * st.ys.v[k];
        }
        double r_r = wdot_partial(r, r), q = wdot_partial(wv, pt.x);
        wsum2(r_r, q);                                         // :56
        const double norm_r0 = sqrt(r_r);
        double z_r = r_r, d_Pd = r_r, e_Pe = 0.0, e_Pd = 0.0, model_value = 0.0;
        TcgResult res;
        res.stop = RIPTRM_TCG_MAX_INNER_ITER;                  // :95
        const int maxinner = o.tcg_maxinner < 0 ? dim(ctx) : o.tcg_maxinner;
        const double Delta2 = Delta * Delta;
        const double nr_theta = (o.tcg_theta == 1.0) ? norm_r0 : pow(norm_r0, o.tcg_theta);
        const double target = norm_r0 * fmin(nr_theta, o.tcg_kappa);
        const double target_sq = target * target;
        double inv_zr = 1.0 / z_r;
        int j = 0;
        for (; j < maxinner; ++j) {                            // :98
            const Vec Sv = matvec(ctx, delta);                 // :100
            Vec tmp;
#pragma unroll
            for (int k = 0; k < K; ++k) tmp.v[k] = st.ys.v[k] * delta.v[k];
            double a, b, g1, h1, h2, h3;
            if (RS) {
                double s6[6] = {wdot_partial(pt.x, Sv), wdot_partial(pt.x, delta), wdot_partial(wv, delta),
                                wdot_partial(delta, Sv), wdot_partial(delta, delta), wdot_partial(delta, tmp)};
                wsum_smem<6>(s6, ctx.ws);
                a = s6[0], b = s6[1], g1 = s6[2], h1 = s6[3], h2 = s6[4], h3 = s6[5];
            } else {
                double s6[8] = {wdot_partial(pt.x, Sv), wdot_partial(pt.x, delta), wdot_partial(wv, delta),
                                wdot_partial(delta, Sv), wdot_partial(delta, delta), wdot_partial(delta, tmp), 0.0, 0.0};
                wsum8x<6>(s6);
                a = s6[0], b = s6[1], g1 = s6[2], h1 = s6[3], h2 = s6[4], h3 = s6[5];
            }
            const double d = ctx.embedded ? g1 : (g1 - b * q);
            Vec Hd;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double ga = ctx.embedded ? delta.v[k] : (delta.v[k] - pt.x.v[k] * b);
                const double t = st.ys.v[k] * ga;
                const double hl = (-Sv.v[k] + a * pt.x.v[k]) + st.kappa * delta.v[k];
                const double g = t - d * pt.x.v[k];
                Hd.v[k] = hl + g;
            }
            const double dt = ctx.embedded ? h3 : (h3 - b * g1);
            const double d_Hd = (((-h1 + a * b) + st.kappa * h2) + dt) - d * b;   // :103
            double alpha = 0.0, e_Pe_new = e_Pe;
            if (d_Hd != 0.0) {                                 // :106-114
                alpha = z_r / d_Hd;
                e_Pe_new = (e_Pe + (2.0 * alpha) * e_Pd) + (alpha * alpha) * d_Pd;
            }
            if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {           // :118
                const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;  // :123-125
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    eta.v[k] = eta.v[k] + tau * delta.v[k];    // :127
                    Heta.v[k] = Heta.v[k] + tau * Hd.v[k];     // :132
                }
                res.stop = (d_Hd <= 0.0) ? RIPTRM_TCG_NEGATIVE_CURVATURE : RIPTRM_TCG_EXCEEDED_TR;
                ++j;
                break;
            }
            e_Pe = e_Pe_new;                                   // :149
            Vec new_eta, new_Heta;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                new_eta.v[k] = eta.v[k] + alpha * delta.v[k];  // :150
                new_Heta.v[k] = Heta.v[k] + alpha * Hd.v[k];   // :154
                r.v[k] = r.v[k] + alpha * Hd.v[k];             // :172 (in place: r is not read after the :163 exit)
            }
            double s4[4] = {wdot_partial(new_eta, st.c), wdot_partial(new_eta, new_Heta), wdot_partial(r, r),
                            wdot_partial(pt.x, r)};
            if (RS) wsum_smem<4>(s4, ctx.ws); else wsum4x(s4);
            const double new_model = s4[0] + 0.5 * s4[1];      // :86-87, :162
            if (new_model >= model_value) {                    // :163
                res.stop = RIPTRM_TCG_MODEL_INCREASED;
                ++j;
                break;
            }
            eta = new_eta;                                     // :167-169
            Heta = new_Heta;
            model_value = new_model;
            r_r = s4[2];                                       // :175
            // ||r|| <= target tested on squares: no square root on the critical path
            if (j >= o.tcg_mininner && r_r <= target_sq) {     // :183-191
                res.stop = (o.tcg_kappa < nr_theta) ? RIPTRM_TCG_REACHED_TARGET_LINEAR
                                                    : RIPTRM_TCG_REACHED_TARGET_SUPERLINEAR;
                ++j;
                break;
            }
            // beta = z_r / z_r_old (:205) as a product with the reciprocal formed one iteration earlier: the division
            // for the next iteration overlaps with the next S.delta instead of sitting on the critical path
            const double beta = r_r * inv_zr;
            z_r = r_r;                                         // :200-202
            inv_zr = 1.0 / z_r;
            const double xd = -s4[3] + beta * b;               // <x, -r + beta delta>
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double dn = -r.v[k] + beta * delta.v[k]; // :206
                delta.v[k] = dn - xd * pt.x.v[k];              // :210
            }
            e_Pd = beta * (e_Pd + alpha * d_Pd);               // :213
            d_Pd = z_r + (beta * beta) * d_Pd;                 // :214
        }
        res.iters = j;
        res.model_value = model_value;
        return res;
    }

    // ---- Exact_RepMat: orthonormal tangent basis (solver_warp.cuh `RepWork`; riptrm_b200/basis.py sphere_basis) -------------
    // H = I - beta u u', u = x + sigma e_0, beta = 2 / u'u maps e_0 to -sigma x / |x|, so its columns 1..n-1 are an orthonormal
    // basis of the tangent space x^perp: coordinates of v are rows 1..n-1 of H v, and sum_i coef_i b_i = H (0; coef).
    struct Coord {
        Vec u;
        double beta;
    };
    static __device__ __forceinline__ void coord_setup(const Ctx&, const Pt& pt, Coord& cc) {
        const double x0 = wbcast(pt.x.v[0], 0);                 // element 0 lives on lane 0, slot 0
        cc.u = pt.x;
        if (lane_id() == 0) cc.u.v[0] = pt.x.v[0] + (x0 >= 0.0 ? 1.0 : -1.0);
        cc.beta = 2.0 / wdot(cc.u, cc.u);
    }
    static __device__ __forceinline__ Vec from_coords(const Ctx& c, const Pt&, const Coord& cc, const double* coef) {
        Vec w;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int e = elem(k);
            w.v[k] = (e >= 1 && e < dimn(c)) ? coef[e - 1] : 0.0;
        }
        const double f = cc.beta * wdot(cc.u, w);
#pragma unroll
        for (int k = 0; k < K; ++k) w.v[k] = w.v[k] - f * cc.u.v[k];
        return w;
    }
    static __device__ __forceinline__ void to_coords(const Ctx& c, const Pt&, const Coord& cc, const Vec& v, double* out) {
        const double f = cc.beta * wdot(cc.u, v);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int e = elem(k);
            if (e >= 1 && e < dimn(c)) out[e - 1] = v.v[k] - f * cc.u.v[k];
        }
        __syncwarp();
    }

    static __device__ __forceinline__ Vec retract(const Ctx&, const Pt& pt, const Vec& dx) {
        Vec a;
#pragma unroll
        for (int k = 0; k < K; ++k) a.v[k] = pt.x.v[k] + dx.v[k];
        const double nrm = sqrt(wdot(a, a));
#pragma unroll
        for (int k = 0; k < K; ++k) a.v[k] = a.v[k] / nrm;
        return a;
    }

    // grad f(x) + sum_i y_i grad g_i(x), grad g_i = -P_x(e_i); `xy` = <x, y>
    static __device__ __forceinline__ double gradL_xy_partial(const Ctx&, const Pt& pt, const CVec& y) {
        return wdot_partial(pt.x, y);
    }
    static __device__ __forceinline__ double gradL_norm_given(const Ctx&, const Pt& pt, const CVec& y, double xy) {
        Vec g;
#pragma unroll
        for (int k = 0; k < K; ++k)
            g.v[k] = (-pt.Sx.v[k] + pt.xSx * pt.x.v[k]) - (y.v[k] - xy * pt.x.v[k]);
        return sqrt(wdot(g, g));
    }
    static __device__ __forceinline__ double gradL_norm(const Ctx& c, const Pt& pt, const CVec& y) {
        return gradL_norm_given(c, pt, y, wsum(gradL_xy_partial(c, pt, y)));
    }

    static __device__ __forceinline__ double manvio(const Ctx&, const Pt& pt) {
        return sqrt(wdot(pt.x, pt.x)) - 1.0;  // src/NonnegPCA/simulator.py:12-14
    }

    static __device__ __forceinline__ double dist(const Ctx&, const Vec& xPrev, const Pt& pt) {
        const double ip = fmax(fmin(wdot(xPrev, pt.x), 1.0), -1.0);  // pymanopt Sphere.dist
        return acos(ip);
    }
};

}  // namespace riptrm
