// pcx_kernels.cuh -- hand-written sm_100a kernel skeleton of the pycollo_b200 engine.
//
// Compiled at pcx_create() time by NVRTC (--gpu-architecture=sm_100a) together
// with the generated per-problem header "pcx_problem.h" (device functions for
// the per-node expression bodies + constexpr dimension/offset tables, emitted by
// pycollo_b200/codegen.py).  Everything else -- tiling, shared-memory staging,
// the collocation contractions, the coalesced scatter into the fixed CCS
// pattern, warp-shuffle reductions and the last-CTA border pass -- is below.
//
// One CTA = one tile = a contiguous range of mesh sections of one phase
// (pycollo_b200/structure.py).  One launch evaluates any subset of
//   C : constraint vector            (pycollo/backend.py:1513-1672)
//   DY: state derivatives            (backend.py:1541-1549, 1665-1668)
//   G : Jacobian non-zeros, CCS      (backend.py:1674-1679, 1747-1761)
//   H : Lagrangian Hessian, triu CCS (backend.py:1693, nlpsol convention)
//   J / GRAD: objective and gradient (backend.py:1495-1511)
// selected at compile time by PCX_FLAGS so unused arithmetic is eliminated.
//
// Bound: HBM streaming.  Algorithmic bytes per evaluation are 8*(num_x+nnz_G)
// for G and 8*(num_x+num_c+nnz_H) for H; the integer
// pattern is never re-read at full size: section descriptors are 20 B per mesh
// section and the recipe table is per section *type* (L1/L2 resident).

#include "pcx_problem.h"

#include "pcx_params.h"

#ifndef PCX_FLAGS
#define PCX_FLAGS (PCX_F_G | PCX_F_H)
#endif
#ifndef PCX_THREADS
#define PCX_THREADS 128
#endif

__device__ __forceinline__ double pcx_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed shuffle tree + fixed warp order).
__device__ __forceinline__ double pcx_block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = pcx_warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < PCX_THREADS / 32; ++w) r += scratch[w];
    }
    return r;   // valid on thread 0
}

template <int N> struct PcxArr { double v[N > 0 ? N : 1]; };

// x = V * x_tilde + r exactly as the reference composes it (a rounded product,
// then a rounded sum -- pycollo/backend.py:279-280, scaling.py:176-178): no FMA
// contraction, so e.g. 1000 * (-0.4) + 500 is exactly 100 as it is in CasADi.
__device__ __forceinline__ double pcx_unscale(double V, double xt, double r) {
    return __dadd_rn(__dmul_rn(V, xt), r);
}

// ---------------------------------------------------------------------------
// One tile of phase Ph.
// ---------------------------------------------------------------------------
template <class Ph>
__device__ void pcx_tile(const PcxParams& p, const int tile, const int inst,
                         unsigned char* smem_raw)
{
    constexpr int F = PCX_FLAGS;
    constexpr int NY = Ph::NY, NV = Ph::NV, NP = Ph::NP, NQ = Ph::NQ, NF = Ph::NF;
    constexpr int NS = PCX_NS;
    constexpr int T = PCX_THREADS;
    constexpr bool WANT_C = (F & PCX_F_C) != 0, WANT_DY = (F & PCX_F_DY) != 0;
    constexpr bool WANT_G = (F & PCX_F_G) != 0, WANT_H = (F & PCX_F_H) != 0;
    constexpr bool WANT_GRAD = (F & PCX_F_GRAD) != 0;
    constexpr bool HAS_T = Ph::HAS_T0 || Ph::HAS_TF;
    // row-contraction sources: f values (C, or G with free time) and the
    // s-derivatives of state equations (G)
    constexpr bool NEED_SF = WANT_C || (WANT_G && HAS_T);
    constexpr int NDS = Ph::ND1SD;                 // d-family s-derivative entries
    constexpr bool NEED_ROWS = NEED_SF || (WANT_G && NDS > 0);

    const int tid = threadIdx.x;
    const i64* pb = p.pbase + Ph::PBASE_OFF;
    const double* ps = p.pscal + Ph::PSCAL_OFF;
    const i64 N = pb[Ph::PB_N], K = pb[Ph::PB_K];
    const i64 xo = pb[Ph::PB_XOFF], co = pb[Ph::PB_COFF];
    const i64 sec_off = pb[Ph::PB_SECOFF];         // offset into sec_order/h/type
    const i64* sec_node = p.sec_node + sec_off + Ph::INDEX;   // K+1 per phase
    const int k0 = p.tile_k0[tile], k1 = p.tile_k1[tile];
    const int nsec = k1 - k0;
    const i64 node0 = sec_node[k0];
    const int nn = (int)(sec_node[k1] - node0) + 1;
    const bool last_tile = (k1 == (int)K);
    const bool has_prev = (k0 > 0);
    const int nnp = nn | 1;                        // odd stride: no bank conflicts

    const double* x = p.x + (i64)inst * p.num_x;
    const double* lam = WANT_H ? p.lam + (i64)inst * p.num_c : nullptr;

    // ---- shared memory carve-up ------------------------------------------
    double* sB = reinterpret_cast<double*>(smem_raw);          // btab
    double* sHk = sB + p.btab_len;                             // nsec+1 (prev first)
    double* sF = sHk + (nsec + 1);                             // NY * nnp
    double* sD = sF + (NEED_SF ? NY * nnp : 0);                // ND1V * nnp
    double* sDS = sD + (WANT_G ? Ph::ND1V * nnp : 0);          // NDS * nnp
    double* sLam = sDS + (WANT_G ? NDS * nnp : 0);             // NY * (nn + 16)
    const int lam_stride = nn + 16;
    double* sRed = sLam + (WANT_H ? NY * lam_stride : 0);      // T/32
    int* sSecNode = reinterpret_cast<int*>(sRed + T / 32);     // nsec+2 (prev first)
    int* sSecOrder = sSecNode + (nsec + 2);                    // nsec+1 (prev first)
    int* sSecType = sSecOrder + (nsec + 1);                    // nsec
    int* sNodeSec = sSecType + nsec;                           // nn
    int* sStart = sNodeSec + nn;                               // NV * (nsec+1)
    int* sRecBase = sStart + NV * (nsec + 1);                  // NV * nsec

    for (int i = tid; i < p.btab_len; i += T) sB[i] = p.btab[i];
    for (int s = tid; s <= nsec; s += T) {
        const int k = k0 - 1 + s;                  // s = 0 is the previous section
        const bool ok = (k >= 0);
        sHk[s] = ok ? p.sec_h[sec_off + k] : 0.0;
        sSecOrder[s] = ok ? p.sec_order[sec_off + k] : 0;
        sSecNode[s] = ok ? (int)(sec_node[k] - node0) : 0;
        if (s > 0) sSecType[s - 1] = p.sec_type[sec_off + k];
    }
    if (tid == 0) sSecNode[nsec + 1] = nn - 1;
    __syncthreads();
    for (int s = tid; s < nsec; s += T) {
        const int b = sSecNode[s + 1], n = sSecOrder[s + 1];
        for (int m = 0; m < n - 1; ++m) sNodeSec[b + m] = s;
        if (s == nsec - 1) sNodeSec[b + n - 1] = s;
    }
    if (WANT_G) {
        const bool uni = p.tile_uniform[tile] != 0;
        const i64* gp = p.gsec_ptr + pb[Ph::PB_GSECOFF];
        for (int i = tid; i < NV * (nsec + 1); i += T) {
            const int a = i / (nsec + 1), s = i - a * (nsec + 1);
            int v;
            if (uni) {
                const int ty = p.sec_type[sec_off + k0];
                const int* tv = p.type_var_off + ty * (p.nvmax + 1);
                v = s * (tv[a + 1] - tv[a]);
            } else {
                v = (int)(gp[a * (K + 1) + k0 + s] - gp[a * (K + 1) + k0]);
            }
            sStart[i] = v;
        }
        for (int i = tid; i < NV * nsec; i += T) {
            const int a = i / nsec, s = i - a * nsec;
            sRecBase[i] = p.type_var_off[p.sec_type[sec_off + k0 + s] * (p.nvmax + 1) + a];
        }
    }
    if (WANT_H) {
        // multipliers of the defect rows of sections k0-1 .. k1-1
        const int prev_rows = has_prev ? sSecOrder[0] - 1 : 0;
        const int nrows = nn - 1 + prev_rows;
        const i64 row0 = node0 - prev_rows;
        for (int i = tid; i < NY * nrows; i += T) {
            const int st = i / nrows, r = i - st * nrows;
            sLam[st * lam_stride + r] = lam[co + (i64)st * (N - 1) + row0 + r];
        }
    }
    __syncthreads();

    // ---- phase scalars ------------------------------------------------------
    const double t0 = pcx_unscale(ps[Ph::OFF_TINFO + 0],
                                  Ph::HAS_T0 ? x[pb[Ph::PB_T0X]] : 0.0, ps[Ph::OFF_TINFO + 1]);
    const double tF = pcx_unscale(ps[Ph::OFF_TINFO + 2],
                                  Ph::HAS_TF ? x[pb[Ph::PB_TFX]] : 0.0, ps[Ph::OFF_TINFO + 3]);
    const double hp = 0.5 * (tF - t0);
    double sv[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < NS; ++j)
        sv[j] = pcx_unscale(p.gscal[PCX_GS_VS + j], x[p.num_x - NS + j], p.gscal[PCX_GS_RS + j]);

    double red[Ph::NRED > 0 ? Ph::NRED : 1];
#pragma unroll
    for (int k = 0; k < Ph::NRED; ++k) red[k] = 0.0;

    double* out_c = WANT_C ? p.c + (i64)inst * p.num_c : nullptr;
    double* out_g = WANT_G ? p.gj + (i64)inst * p.nnz_g : nullptr;
    double* out_h = WANT_H ? p.hs + (i64)inst * p.nnz_h : nullptr;
    double* out_dy = WANT_DY ? p.dy + (i64)inst * p.num_dy : nullptr;
    double* out_grad = WANT_GRAD ? p.grad + (i64)inst * p.num_x : nullptr;
    double* bv = p.bv + (i64)inst * p.bv_size;

    // ---- node-parallel evaluation ---------------------------------------------
    for (int ml = tid; ml < nn; ml += T) {
        const bool owned = (ml < nn - 1) || last_tile;
        const int s = sNodeSec[ml];
        const int mloc = ml - sSecNode[s + 1];
        const int n_k = sSecOrder[s + 1];
        const double h_k = sHk[s + 1];
        const i64 m = node0 + ml;
        const bool start_with_prev = (mloc == 0) && (s > 0 || has_prev);
        const int n_pr = start_with_prev ? sSecOrder[s] : 0;
        const double h_pr = start_with_prev ? sHk[s] : 0.0;

        double v[NV + NS > 0 ? NV + NS : 1];
#pragma unroll
        for (int a = 0; a < NV; ++a)
            v[a] = pcx_unscale(ps[Ph::OFF_VV + a], x[xo + (i64)a * N + m], ps[Ph::OFF_RV + a]);
#pragma unroll
        for (int j = 0; j < NS; ++j) v[NV + j] = sv[j];

        // quadrature weight of this node, accumulated as pycollo/mesh.py:325-326
        double wq = 0.0;
        if (NQ > 0 && (WANT_C || WANT_G || WANT_H)) {
            if (start_with_prev)
                wq = __dmul_rn(sB[p.order_w_off[n_pr] + n_pr - 1], h_pr);
            wq = __dadd_rn(wq, __dmul_rn(sB[p.order_w_off[n_k] + mloc], h_k));
        }

        double muh[NF > 0 ? NF : 1], mut[NF > 0 ? NF : 1];
#pragma unroll
        for (int e = 0; e < NF; ++e) { muh[e] = 0.0; mut[e] = 0.0; }
        if (WANT_H) {
            const int prev_rows = has_prev ? sSecOrder[0] - 1 : 0;
#pragma unroll
            for (int i = 0; i < NY; ++i) {
                double acc = 0.0;
                const double* lrow = sLam + i * lam_stride + prev_rows;
                if (start_with_prev) {
                    const double* Apr = sB + p.order_a_off[n_pr];
                    const int rb = sSecNode[s];         // first row of prev section
                    for (int l = 0; l < n_pr - 1; ++l)
                        acc += lrow[rb + l] * __dmul_rn(Apr[l * n_pr + n_pr - 1], h_pr);
                }
                if (owned) {
                    const double* Ak = sB + p.order_a_off[n_k];
                    const int rb = sSecNode[s + 1];
                    for (int l = 0; l < n_k - 1; ++l)
                        acc += lrow[rb + l] * __dmul_rn(Ak[l * n_k + mloc], h_k);
                }
                const double mu = ps[Ph::OFF_WFN + i] * acc;
                mut[i] = mu;
                muh[i] = hp * mu;
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const double mu = owned ? ps[Ph::OFF_WFN + NY + j]
                    * lam[co + (i64)NY * (N - 1) + (i64)j * N + m] : 0.0;
                muh[NY + j] = mu;
                mut[NY + j] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const double mu = -ps[Ph::OFF_WFN + NY + NP + i]
                    * lam[co + (i64)NY * (N - 1) + (i64)NP * N + i] * wq;
                mut[NY + NP + i] = mu;
                muh[NY + NP + i] = hp * mu;
            }
        }

        PcxArr<NF> Fv; PcxArr<Ph::ND1V> D1V; PcxArr<Ph::ND1S> D1S;
        PcxArr<Ph::NH2VV> H2VV; PcxArr<Ph::NH2VS> H2VS; PcxArr<Ph::NH2SS> H2SS;
        PcxArr<Ph::NHTV> HTV; PcxArr<Ph::NHTS> HTS;
        Ph::eval(v, muh, mut, Fv.v, D1V.v, D1S.v, H2VV.v, H2VS.v, H2SS.v, HTV.v, HTS.v);

        if (NEED_SF) {
#pragma unroll
            for (int i = 0; i < NY; ++i) sF[i * nnp + ml] = Fv.v[i];
        }
        if (WANT_DY && owned) {
#pragma unroll
            for (int i = 0; i < NY; ++i)
                out_dy[pb[Ph::PB_DYOFF] + (i64)i * N + m] = Fv.v[i];
        }
        if (WANT_C && owned) {
#pragma unroll
            for (int j = 0; j < NP; ++j)
                out_c[co + (i64)NY * (N - 1) + (i64)j * N + m] =
                    ps[Ph::OFF_WFN + NY + j] * Fv.v[NY + j];
        }
        if ((WANT_C || (WANT_G && HAS_T)) && owned) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) red[Ph::RED_G + i] += wq * Fv.v[NY + NP + i];
        }
        if (WANT_G) {
#pragma unroll
            for (int k = 0; k < Ph::ND1V; ++k) {
                const int fam = Ph::FAM(Ph::D1V_FN(k));
                const double fac = fam == 0 ? hp : (fam == 1 ? 1.0 : -hp * wq);
                sD[k * nnp + ml] = ps[Ph::OFF_D1V + k] * fac * D1V.v[k];
            }
            int kd = 0, kr = 0;
#pragma unroll
            for (int k = 0; k < Ph::ND1S; ++k) {
                const int fam = Ph::FAM(Ph::D1S_FN(k));
                if (fam == 0) {
                    sDS[kd * nnp + ml] = ps[Ph::OFF_D1S + k] * hp * D1S.v[k];
                    ++kd;
                } else if (fam == 1) {
                    if (owned) out_g[pb[Ph::PB_GSCOL + k] + m] =
                        ps[Ph::OFF_D1S + k] * D1S.v[k];
                } else {
                    if (owned) red[Ph::RED_GS + kr] += wq * D1S.v[k];
                    ++kr;
                }
            }
        }
        if (WANT_H && owned) {
            const bool irregular = (m == 0) || (m == N - 1);
            if (!irregular) {
#pragma unroll
                for (int k = 0; k < Ph::NH2VV; ++k)
                    out_h[pb[Ph::PB_HREG + Ph::H2VV_B(k)]
                          + (m - 1) * Ph::NA(Ph::H2VV_B(k)) + Ph::H2VV_POS(k)] =
                        ps[Ph::OFF_H2VV + k] * H2VV.v[k];
#pragma unroll
                for (int k = 0; k < Ph::NH2VS; ++k)
                    out_h[pb[Ph::PB_HS + k] + (m - 1)] = ps[Ph::OFF_H2VS + k] * H2VS.v[k];
#pragma unroll
                for (int k = 0; k < Ph::NHTV; ++k) {
                    if (Ph::HAS_T0)
                        out_h[pb[Ph::PB_HT0 + k] + (m - 1)] = ps[Ph::OFF_HT0 + k] * HTV.v[k];
                    if (Ph::HAS_TF)
                        out_h[pb[Ph::PB_HTF + k] + (m - 1)] = ps[Ph::OFF_HTF + k] * HTV.v[k];
                }
            } else {
                double* irr = bv + pb[m == 0 ? Ph::PB_IRR0 : Ph::PB_IRR1];
#pragma unroll
                for (int k = 0; k < Ph::NH2VV; ++k) irr[k] = H2VV.v[k];
#pragma unroll
                for (int k = 0; k < Ph::NH2VS; ++k) irr[Ph::NH2VV + k] = H2VS.v[k];
#pragma unroll
                for (int k = 0; k < Ph::NHTV; ++k) irr[Ph::NH2VV + Ph::NH2VS + k] = HTV.v[k];
            }
#pragma unroll
            for (int k = 0; k < Ph::NHTS; ++k) red[Ph::RED_HTS + k] += HTS.v[k];
#pragma unroll
            for (int k = 0; k < Ph::NH2SS; ++k) red[Ph::RED_HSS + k] += H2SS.v[k];
        }
        if (WANT_GRAD && owned) {
#pragma unroll
            for (int a = 0; a < NV; ++a) out_grad[xo + (i64)a * N + m] = 0.0;
        }
    }
    if (WANT_GRAD && tile == 0) {
        // non-node entries (q, t of every phase, s) are zeroed once; the border
        // pass overwrites the structural ones afterwards
        for (int i = tid; i < PCX_NPOINT; i += T) out_grad[p.pt_x[i]] = 0.0;
    }
    __syncthreads();

    // ---- row-oriented contractions: defect rows of c, t/s columns of G -------
    if (NEED_ROWS) {
        for (int r = tid; r < nn - 1; r += T) {
            // row r belongs to the section that contains node r+1 as a
            // non-start node: that is the section of node r
            const int s = sNodeSec[r];
            const int b = sSecNode[s + 1];
            const int l = r - b;
            const int n_k = sSecOrder[s + 1];
            const double h_k = sHk[s + 1];
            const double* Arow = sB + p.order_a_off[n_k] + l * n_k;
            if (NEED_SF) {
#pragma unroll
                for (int i = 0; i < NY; ++i) {
                    double acc = 0.0;
                    for (int mm = 0; mm < n_k; ++mm)
                        acc += __dmul_rn(Arow[mm], h_k) * sF[i * nnp + b + mm];
                    if (WANT_C) {
                        const double ya = pcx_unscale(ps[Ph::OFF_VV + i],
                            x[xo + (i64)i * N + node0 + b], ps[Ph::OFF_RV + i]);
                        const double yb = pcx_unscale(ps[Ph::OFF_VV + i],
                            x[xo + (i64)i * N + node0 + r + 1], ps[Ph::OFF_RV + i]);
                        out_c[co + (i64)i * (N - 1) + node0 + r] =
                            ps[Ph::OFF_WFN + i] * ((ya - yb) + hp * acc);
                    }
                    if (WANT_G && Ph::HAS_T0 && Ph::FN_NZ(i))
                        out_g[pb[Ph::PB_GT0 + i] + node0 + r] = ps[Ph::OFF_GT0 + i] * acc;
                    if (WANT_G && Ph::HAS_TF && Ph::FN_NZ(i))
                        out_g[pb[Ph::PB_GTF + i] + node0 + r] = ps[Ph::OFF_GTF + i] * acc;
                }
            }
            if (WANT_G && NDS > 0) {
                int kd = 0;
#pragma unroll
                for (int k = 0; k < Ph::ND1S; ++k) {
                    if (Ph::FAM(Ph::D1S_FN(k)) == 0) {
                        double acc = 0.0;
                        for (int mm = 0; mm < n_k; ++mm)
                            acc += __dmul_rn(Arow[mm], h_k) * sDS[kd * nnp + b + mm];
                        out_g[pb[Ph::PB_GSCOL + k] + node0 + r] = acc;
                        ++kd;
                    }
                }
            }
        }
    }

    // ---- coalesced scatter of the Jacobian values, one variable at a time ----
    if (WANT_G) {
#pragma unroll 1
        for (int a = 0; a < NV; ++a) {
            const int* st = sStart + a * (nsec + 1);
            const int len = st[nsec];
            if (len == 0) continue;
            const i64 base = p.tile_gbase[(i64)tile * p.nvmax + a];
            const float inv = (float)nsec / (float)len;
            const double* cst = ps + Ph::OFF_GCST;
            for (int idx = tid; idx < len; idx += T) {
                int s = min(nsec - 1, (int)((float)idx * inv));
                while (idx < st[s]) --s;
                while (idx >= st[s + 1]) ++s;
                const int local = idx - st[s];
                const u32 w = __ldg(p.recipes + sRecBase[a * nsec + s] + local);
                if (w >> RC_SKIP_BIT) continue;
                const int e = w & ((1u << RC_E_BITS) - 1);
                const int bi = (w >> RC_B_SHIFT) & ((1u << RC_B_BITS) - 1);
                const int ml = sSecNode[s + 1] + (int)((w >> RC_M_SHIFT) & ((1u << RC_M_BITS) - 1));
                const int ci = (w >> RC_C_SHIFT) & ((1u << RC_C_BITS) - 1);
                const bool prev = (w >> RC_PREV_BIT) & 1u;
                const bool plain = (w >> RC_PLAIN_BIT) & 1u;
                const double d = e ? sD[(e - 1) * nnp + ml] : 0.0;
                const double coef = plain ? 1.0 : __dmul_rn(sB[bi], prev ? sHk[s] : sHk[s + 1]);
                out_g[base + idx] = coef * d + cst[ci];
            }
        }
    }

    // ---- reductions -> per-tile partials ---------------------------------------
    if (Ph::NRED > 0) {
#pragma unroll
        for (int k = 0; k < Ph::NRED; ++k) {
            const double r = pcx_block_sum(red[k], sRed);
            if (tid == 0)
                p.partials[((i64)inst * p.num_tiles + tile) * p.nred_max + k] = r;
        }
    }
}

// ---------------------------------------------------------------------------
// Border pass: executed by the last CTA to finish (per instance).
// ---------------------------------------------------------------------------
__device__ void pcx_border(const PcxParams& p, const int inst, double* scratch)
{
    constexpr int F = PCX_FLAGS;
    constexpr int T = PCX_THREADS;
    const int tid = threadIdx.x;
    double* bv = p.bv + (i64)inst * p.bv_size;
    const double* x = p.x + (i64)inst * p.num_x;
    __shared__ double sRS[1 + PCX_NUM_PHASES];

    // 1. final reductions: deterministic (fixed stride, fixed shuffle tree)
    for (int q = 0; q < PCX_NUM_PHASES; ++q) {
        const int nred = PCX_PHASE_NRED(q);
        const i64* pbq = p.pbase + PCX_PHASE_PBASE(q);
        const int t_lo = (int)pbq[PCX_PB_TILE0], t_hi = (int)pbq[PCX_PB_TILE1];
        for (int k = 0; k < nred; ++k) {
            double acc = 0.0;
            for (int t = t_lo + tid; t < t_hi; t += T)
                acc += p.partials[((i64)inst * p.num_tiles + t) * p.nred_max + k];
            const double r = pcx_block_sum(acc, scratch);
            if (tid == 0) bv[PCX_PHASE_REDOFF(q) + k] = r;
        }
    }
    // 2. point functions and runtime scalars
    if (tid == 0) {
        bv[0] = 1.0;
        double pt[PCX_NPOINT > 0 ? PCX_NPOINT : 1];
        for (int a = 0; a < PCX_NPOINT; ++a) {
            pt[a] = pcx_unscale(p.pt_scal[a], x[p.pt_x[a]], p.pt_scal[PCX_NPOINT + a]);
            bv[PCX_BV_PTVAL + a] = pt[a];
        }
        double mult[1 + PCX_NB];
        const double sg = (p.sigma != nullptr) ? p.sigma[inst] : 1.0;
        mult[0] = sg * p.gscal[PCX_GS_W];
        for (int k = 0; k < PCX_NB; ++k)
            mult[1 + k] = (F & PCX_F_H)
                ? p.lam[(i64)inst * p.num_c + (p.num_c - PCX_NB) + k] * p.gscal[PCX_GS_WB + k]
                : 0.0;
        pcx_point_eval(pt, mult, bv + PCX_BV_PTFN, bv + PCX_BV_PTD1, bv + PCX_BV_PTD2);
        sRS[0] = 1.0;
        for (int q = 0; q < PCX_NUM_PHASES; ++q) {
            const i64* pb = p.pbase + PCX_PHASE_PBASE(q);
            const double* ps = p.pscal + PCX_PHASE_PSCAL(q);
            const i64 i0 = pb[PCX_PB_T0X], iF = pb[PCX_PB_TFX];
            const double t0 = pcx_unscale(ps[PCX_PHASE_TINFO(q) + 0], i0 >= 0 ? x[i0] : 0.0,
                                          ps[PCX_PHASE_TINFO(q) + 1]);
            const double tF = pcx_unscale(ps[PCX_PHASE_TINFO(q) + 2], iF >= 0 ? x[iF] : 0.0,
                                          ps[PCX_PHASE_TINFO(q) + 3]);
            sRS[1 + q] = 0.5 * (tF - t0);
        }
    }
    __syncthreads();
    __threadfence();
    // 3. the border map
    for (int e = tid; e < p.n_border; e += T) {
        const int grp = p.border_grp[e];
        double* out;
        if (grp == 0) { if (!(F & PCX_F_C)) continue; out = p.c + (i64)inst * p.num_c; }
        else if (grp == 1) { if (!(F & PCX_F_G)) continue; out = p.gj + (i64)inst * p.nnz_g; }
        else if (grp == 2) { if (!(F & PCX_F_H)) continue; out = p.hs + (i64)inst * p.nnz_h; }
        else if (grp == 3) { if (!(F & PCX_F_J)) continue; out = p.jval + inst; }
        else { if (!(F & PCX_F_GRAD)) continue; out = p.grad + (i64)inst * p.num_x; }
        double acc = 0.0;
        for (int k = p.border_ptr[e]; k < p.border_ptr[e + 1]; ++k)
            acc += p.border_coef[k] * bv[p.border_bv[k]] * sRS[p.border_rs[k]];
        out[p.border_slot[e]] = acc;
    }
}

extern "C" __global__ void __launch_bounds__(PCX_THREADS)
PCX_KERNEL_NAME(const PcxParams p)
{
    extern __shared__ __align__(16) unsigned char pcx_smem[];
    __shared__ int sLast;
    const int tile = blockIdx.x, inst = blockIdx.y;
    const int phase = p.tile_phase[tile];
    switch (phase) {
#define PCX_CASE(P) case P: pcx_tile<PcxPhase<P> >(p, tile, inst, pcx_smem); break;
        PCX_FOREACH_PHASE(PCX_CASE)
#undef PCX_CASE
        default: break;
    }
    // ticket: the last CTA of this instance runs the border pass
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const u32 t = atomicAdd(p.ticket + inst, 1u);
        sLast = (t == (u32)p.num_tiles - 1u);
    }
    __syncthreads();
    if (sLast) {
        __threadfence();
        pcx_border(p, inst, reinterpret_cast<double*>(pcx_smem));
        if (threadIdx.x == 0) p.ticket[inst] = 0u;
    }
}
