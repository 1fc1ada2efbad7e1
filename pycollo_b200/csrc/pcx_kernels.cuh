// pcx_kernels.cuh -- hand-written sm_100a kernel skeleton of the pycollo_b200 engine.
//
// Compiled at pcx_create() time by NVRTC (--gpu-architecture=sm_100a) together
// with the generated per-problem header "pcx_problem.h" (device functions for
// the per-node expression bodies + constexpr dimension/offset tables, emitted by
// pycollo_b200/codegen.py).  Everything else -- tiling, shared-memory staging,
// the collocation contractions, the coalesced scatter into the fixed CCS
// pattern, warp-shuffle reductions, the dedicated border CTA, the programmatic
// dependent launch protocol and the multi-GPU border exchange over peer memory
// -- is below.
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
#ifndef PCX_MIN_BLOCKS
#define PCX_MIN_BLOCKS 1
#endif
#ifndef PCX_SCATTER_UNROLL
#define PCX_SCATTER_UNROLL 4
#endif
// Two-pass node phase for large bodies (fused G+H launches): pass 1 evaluates the
// first-derivative outputs, the Jacobian is scattered, then pass 2 evaluates the
// second-derivative outputs INTO THE SHARED MEMORY THE STAGED FIRST DERIVATIVES
// JUST VACATED and the Hessian entries leave as full sectors.  Written directly
// (one 8-byte piece of a different sector per lane) the Hessian stores cost as many
// SM->L2 sector slots as the whole Jacobian (profiles/r02_d3_knockout.txt).
// On by default for bodies with at least PCX_TWO_PASS_MIN node-diagonal Hessian
// entries per node (Delta III 37: 0.468 -> 0.433 ms at 10^6 nodes; shuttle 24:
// 0.381 -> 0.353 ms); -DPCX_TWO_PASS=0 restores the fused single pass.
#ifndef PCX_TWO_PASS
#define PCX_TWO_PASS 1
#endif
#ifndef PCX_TWO_PASS_MIN
#define PCX_TWO_PASS_MIN 16
#endif
#ifndef PCX_STAGGER_NS
#define PCX_STAGGER_NS 0
#endif
#ifndef PCX_INTERLEAVE_PHASES
#define PCX_INTERLEAVE_PHASES 0
#endif
// timing experiments only (wrong results): keep the arithmetic, drop the stores
#ifndef PCX_KO_HST
#define PCX_KO_HST 0
#endif
#ifndef PCX_KO_GST
#define PCX_KO_GST 0
#endif
__device__ double pcx_ko_sink;
#ifndef PCX_DECODE_V1
#define PCX_DECODE_V1 0
#endif
#ifndef PCX_EARLY_WAIT
#define PCX_EARLY_WAIT 0
#endif
// the thread's first scatter work item decoded from the STAGED run tables (one
// dependent global load instead of four); needs the staged decode.  On by default:
// neutral where the dependency wait hides the prologue, +2-3 % on sweeps of independent
// evaluations (headline: 106.0 -> 109.5 k evals/s at 200 steps, 98.6 -> 100.7 k at 20)
#ifndef PCX_PRE_STAGED
#define PCX_PRE_STAGED (!PCX_DECODE_V1)
#endif
#if PCX_PRE_STAGED && PCX_DECODE_V1
#error "PCX_PRE_STAGED needs the staged decode (PCX_DECODE_V1=0)"
#endif
#define PCX_STR2(x) #x
#define PCX_STR(x) PCX_STR2(x)

// Small per-problem tables live in constant memory (set by the host after the
// module is loaded and on every pcx_set_scaling): scaling products, slot bases.
// With compile-time offsets they become direct c[bank][off] operands.
__constant__ double pcx_c_pscal[PCX_PSCAL_TOTAL > 0 ? PCX_PSCAL_TOTAL : 1];
__constant__ double pcx_c_gscal[PCX_GSCAL_TOTAL > 0 ? PCX_GSCAL_TOTAL : 1];
__constant__ long long pcx_c_pbase[PCX_PBASE_TOTAL > 0 ? PCX_PBASE_TOTAL : 1];

// Programmatic dependent launch (griddepcontrol, sm_90+): when the host launches
// with programmatic stream serialisation, the CTAs of this kernel may become
// resident and run their table-only prologue while the previous kernel in the
// stream drains; pcx_grid_dependency_wait() returns once that kernel has
// completed and its writes are visible.  Without the launch attribute both are
// no-ops.
__device__ __forceinline__ void pcx_grid_dependency_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pcx_grid_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

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

#ifdef PCX_DEBUG_TIMELINE
// per-CTA phase timestamps (ns) into the partials scratch, 16 slots per tile
__device__ __forceinline__ void pcx_stamp(const PcxParams& p, int tile, int k) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.partials[(i64)tile * 16 + k] = (double)(t & ((1ull << 40) - 1));
    }
}
#define PCX_STAMP(k) pcx_stamp(p, tile, k)
#define PCX_STAMP_B(k) pcx_stamp(p, 0, k)
__device__ __forceinline__ void pcx_stamp_smid(const PcxParams& p, int tile) {
    if (threadIdx.x == 0) {
        unsigned id;
        asm volatile("mov.u32 %0, %smid;" : "=r"(id));
        p.partials[(i64)tile * 16 + 10] = (double)id;
    }
}
#else
#define PCX_STAMP(k)
#define PCX_STAMP_B(k)
#endif

// Engine tables (tile/section descriptors, recipes) are re-read by every launch
// while tens of MB of values stream through L2: load them with an evict_last
// policy so they stay L2-resident between launches.
__device__ __forceinline__ unsigned long long pcx_policy_keep() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ long long pcx_ld_keep(const long long* ptr, unsigned long long pol) {
    long long v;
    asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ unsigned long long pcx_ld_keep(const unsigned long long* ptr,
                                                          unsigned long long pol) {
    unsigned long long v;
    asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ double pcx_ld_keep(const double* ptr, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ int pcx_ld_keep(const int* ptr, unsigned long long pol) {
    int v;
    asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
    return v;
}

// a load the compiler keeps where it is written (used to put the iterate in flight early)
__device__ __forceinline__ double pcx_ld_f64(const double* ptr) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
    return v;
}

// x = V * x_tilde + r exactly as the reference composes it (a rounded product,
// then a rounded sum -- pycollo/backend.py:279-280, scaling.py:176-178): no FMA
// contraction, so e.g. 1000 * (-0.4) + 500 is exactly 100 as it is in CasADi.
__device__ __forceinline__ double pcx_unscale(double V, double xt, double r) {
    return __dadd_rn(__dmul_rn(V, xt), r);
}

// Multi-wave grids of equal tiles: every CTA of the first wave starts its data phase at
// the same instant (when the previous kernel completes) and takes equally long, so the
// whole GPU moves in lock-step -- all resident tiles in their store phase at once (a
// burst the memory system cannot absorb: the queue makes them FINISH together again),
// all computing at once (memory idle).  Spreading the first wave's start times
// uniformly over one tile lifetime, AFTER the dependency wait, makes the store demand
// smooth; below capacity nothing queues, lifetimes stay equal and the phases stay spread.
__device__ __forceinline__ void pcx_first_wave_stagger() {
#if PCX_STAGGER_NS > 0
    unsigned nsm;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const unsigned first = nsm * PCX_MIN_BLOCKS;
    if (blockIdx.x < first && blockIdx.x > 0) {
        if (threadIdx.x == 0) {
            // CTAs that share an SM have neighbouring indices (the block scheduler fills
            // an SM before it moves on) or indices one SM count apart: either way
            // i mod R differs between them, so that is the coarse part of the delay
            const unsigned i = blockIdx.x, R = PCX_MIN_BLOCKS;
            const unsigned long long wait_ns =
                (unsigned long long)((i % R) * (first / R) + i / R) * PCX_STAGGER_NS / first;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            do {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            } while (t1 - t0 < wait_ns);
        }
        __syncthreads();
    }
#endif
}

// The tiles that share an SM take turns in their Jacobian scatter (two-pass form).
// Resident CTAs of a multi-wave grid of equal tiles run in lock-step (same start, same
// duration: 77 % of Delta III's CTAs start within 1 us of another CTA of their SM), so
// all of them scatter at once -- each at a third of the SM's SM->L2 store rate -- and
// all compute at once with the store path idle.  A per-SM token serialises the
// scatters; the waiting order it creates de-phases the tiles for the rest of the
// launch (24 %; scatter phase 7.0 -> 4.8 us, profiles/r02_d3_lockstep.txt).  The
// holder never waits for another CTA, so the token cannot deadlock; a stale one
// (aborted launch) is stepped over after 200 us.
#ifndef PCX_STORE_TOKEN
#define PCX_STORE_TOKEN 1
#endif
__device__ unsigned pcx_sm_token[1024];
// returns the token's index (thread 0), to be handed to the release
__device__ __forceinline__ unsigned pcx_store_token_acquire(const int tid) {
    unsigned smid = 0;
    if (tid == 0) {
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        smid &= 1023u;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (atomicCAS(&pcx_sm_token[smid], 0u, 1u) != 0u) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 200000ull) break;
        }
    }
    __syncthreads();
    return smid;
}
__device__ __forceinline__ void pcx_store_token_release(const int tid, const unsigned smid) {
    if (tid == 0) atomicExch(&pcx_sm_token[smid], 0u);
}

// Does this (phase, output selection) accumulate any cross-tile reduction?
template <class Ph>
__host__ __device__ constexpr bool pcx_need_red() {
    return (((PCX_FLAGS & PCX_F_C) != 0) && Ph::NQ > 0)
        || (((PCX_FLAGS & PCX_F_G) != 0) && (Ph::HAS_T0 || Ph::HAS_TF) && Ph::NQ > 0)
        || (((PCX_FLAGS & PCX_F_G) != 0) && (Ph::RED_HTS - Ph::RED_GS) > 0)
        || (((PCX_FLAGS & PCX_F_H) != 0) && (Ph::NHTS > 0 || Ph::NH2SS > 0));
}

// ---------------------------------------------------------------------------
// Scatter of the Jacobian values: run setup, recipe decoding, store loop.
// A section of a given type owns, for every variable a, a fixed "period" of P_a
// value slots; the host lists the tile as runs of consecutive same-type sections
// (one run per tile on a uniform mesh, plus a one-section run in the first and
// last tile).  Thread t takes ONE slot u of the concatenated period (all
// variables): its 64-bit recipe word is decoded once per run and the loop over
// the run's sections is LDS, DFMA, STG + two pointer bumps.  Consecutive threads
// write consecutive addresses inside a variable's period, and a variable's
// periods of consecutive sections are back to back.
// ---------------------------------------------------------------------------
struct PcxRun { int s_lo, s_hi, rec0, Ptot, per, u0, sc0, nstep; const int* tv; };
struct PcxSlot { double bcoef, cc; i64 o; int dp, dstep, ostep, cnt; };

template <class Ph>
__device__ __forceinline__ bool pcx_run_setup(const PcxParams& p, const int run, const int tid,
                                              const unsigned long long keep,
                                              const int* sSecOrder, PcxRun& R) {
    constexpr int T = PCX_THREADS;
    R.s_lo = pcx_ld_keep(p.run_slo + run, keep);
    R.s_hi = pcx_ld_keep(p.run_shi + run, keep);
    R.tv = p.type_var_off + pcx_ld_keep(p.run_type + run, keep) * (p.nvmax + 1);
    R.rec0 = R.tv[0];
    R.Ptot = R.tv[Ph::NV] - R.rec0;                  // slots per section, all variables
    if (R.Ptot == 0) return false;
    R.per = 1; R.u0 = tid; R.sc0 = R.s_lo;
    if (R.Ptot < T) {                                // several sections per pass
        R.per = T / R.Ptot;
        const int q = tid / R.Ptot;
        R.u0 = tid - q * R.Ptot;
        R.sc0 = R.s_lo + q;
        if (q >= R.per) return false;
    }
    // all sections of a run share one type, hence one order n_k: the first node
    // of section sc is nd0 + (sc - s_lo) * (n_k - 1), so the staged-derivative
    // pointer advances by a constant (1 for the previous-section copies)
    R.nstep = sSecOrder[R.s_lo + 1] - 1;
    return true;
}

// dp is an offset (in doubles) from the base of the dynamic shared memory (sB);
// constant-only slots read the 0.0 sentinel sB[0] with stride 0
template <class Ph>
__device__ __forceinline__ bool pcx_decode_slot(const PcxParams& p, const int run, const int u,
                                                const PcxRun& R, const unsigned long long keep,
                                                const double* sB, const double* cst,
                                                const int* sSecNode, const int sD_off,
                                                const int sDP_off, const int nnp, const int nsp,
                                                PcxSlot& S) {
    const unsigned long long w = pcx_ld_keep(p.recipes + R.rec0 + u, keep);
    const u32 lo = (u32)w;
    if (lo >> RC_SKIP_BIT) return false;
    const int a = (int)((w >> RC_VAR_SHIFT) & 0xffu);
    const int local = (int)((w >> RC_LOCAL_SHIFT) & ((1u << RC_LOCAL_BITS) - 1));
    const int Pa = R.tv[a + 1] - R.tv[a];
    const int e = lo & ((1u << RC_E_BITS) - 1);
    const int bi = (lo >> RC_B_SHIFT) & ((1u << RC_B_BITS) - 1);
    const int mloc = (int)((w >> RC_M_SHIFT) & ((1u << RC_M_BITS) - 1));
    S.cc = cst[(lo >> RC_C_SHIFT) & ((1u << RC_C_BITS) - 1)];
    const bool prev = (lo >> RC_PREV_BIT) & 1u;
    S.bcoef = sB[bi];                                // 1.0 for plain slots
    S.dp = 0; S.dstep = 0;
    if (e) {
        S.dp = prev ? sDP_off + (e - 1) * nsp + R.sc0
                    : sD_off + (e - 1) * nnp + mloc + sSecNode[R.sc0 + 1];
        S.dstep = prev ? R.per : R.per * R.nstep;
    }
    S.o = pcx_ld_keep(p.run_gbase + (i64)run * p.nvmax + a, keep) + local
          + (i64)(R.sc0 - R.s_lo) * Pa;
    S.ostep = R.per * Pa;
    S.cnt = (R.s_hi - R.sc0 + R.per - 1) / R.per;
    return true;
}

// no divergent branch: every slot kind is  value = bcoef * staged + constant
__device__ __forceinline__ void pcx_store_run(double* o, const int ostep, const double* dp,
                                              const int dstep, const double bcoef,
                                              const double cc, const int cnt) {
#if PCX_KO_GST
    double acc = 0.0;
    _Pragma(PCX_STR(unroll PCX_SCATTER_UNROLL))
    for (int it = 0; it < cnt; ++it) { acc += bcoef * (*dp) + cc; dp += dstep; }
    if (acc == 1.2345e-300) *o = acc;
    return;
#endif
    _Pragma(PCX_STR(unroll PCX_SCATTER_UNROLL))
    for (int it = 0; it < cnt; ++it) {
        *o = bcoef * (*dp) + cc;
        o += ostep; dp += dstep;
    }
}

// ---------------------------------------------------------------------------
// Destination of the per-node results of the generated body PcxPhase<P>::eval.
// Every method is a compile-time-indexed store: nothing the body computes stays
// live after it has been produced (a problem like the Delta III launcher has
// ~90 results per node; held in arrays they alone overflow the register file).
// ---------------------------------------------------------------------------
// MODE 0: every output; 1: first derivatives only; 2: second derivatives only, node-diagonal
// Hessian entries staged in output order; 3: every output, Hessian entries staged likewise
template <class Ph, int MODE = 0>
struct PcxNodeSink {
    static constexpr int F_ = PCX_FLAGS;
    static constexpr bool FIRST = MODE != 2, SECOND = MODE != 1;
    static constexpr bool STG_H = Ph::STAGE_H || MODE >= 2;
    static constexpr int HSTR = MODE == 2 ? Ph::HPS : Ph::HP;
    static constexpr bool WANT_C = (F_ & PCX_F_C) != 0, WANT_DY = (F_ & PCX_F_DY) != 0;
    static constexpr bool WANT_G = (F_ & PCX_F_G) != 0, WANT_H = (F_ & PCX_F_H) != 0;
    static constexpr bool HAS_T = Ph::HAS_T0 || Ph::HAS_TF;
    static constexpr bool NEED_SF = WANT_C || (WANT_G && HAS_T);
    static constexpr int NY = Ph::NY, NP = Ph::NP;

    const double* ps; const i64* pb;
    double *sF, *sD, *sDS, *sDP;          // already offset by the node / section
    double* sH;                           // this node's row of staged Hessian entries
    int hrow;                             // (two-pass form) index among the tile's regular nodes
    int nnp, nsp;
    bool sec_start, owned, regular;
    double hp, wq, h_k, h_pr;
    i64 m, N;
    double *out_c_path, *out_dy, *out_g, *out_h, *irr, *red;

    // rank of first-derivative-by-parameter entry K among those of one family
    static __host__ __device__ constexpr int d1s_rank(int K, int fam) {
        int c = 0;
        for (int k = 0; k < K; ++k) if (Ph::FAM(Ph::D1S_FN(k)) == fam) ++c;
        return c;
    }

    template <int I> __device__ __forceinline__ void F(const double val) const {
        if (!FIRST) return;
        if (I < NY) {
            if (NEED_SF) sF[I * nnp] = val;
            if (WANT_DY && owned) out_dy[(i64)I * N] = val;
        } else if (I < NY + NP) {
            if (WANT_C && owned)
                out_c_path[(i64)(I - NY) * N] = ps[Ph::OFF_WFN + I] * val;
        } else {
            if ((WANT_C || (WANT_G && HAS_T)) && owned)
                red[Ph::RED_G + (I - NY - NP)] += wq * val;
        }
    }
    // defect-family entries are staged already multiplied by the length of the
    // section that owns the row: h_k for rows of the node's own section and, at
    // a section's first node, h_{k-1} for the rows of the previous section
    // (second copy, one slot per section)
    template <int K> __device__ __forceinline__ void D1V(const double val) const {
        if (!WANT_G || !FIRST) return;
        constexpr int fam = Ph::FAM(Ph::D1V_FN(K));
        const double fac = fam == 0 ? hp : (fam == 1 ? 1.0 : -hp * wq);
        const double d = ps[Ph::OFF_D1V + K] * fac * val;
        sD[K * nnp] = fam == 0 ? d * h_k : d;
        if (fam == 0 && sec_start) sDP[K * nsp] = d * h_pr;
    }
    template <int K> __device__ __forceinline__ void D1S(const double val) const {
        if (!WANT_G || !FIRST) return;
        constexpr int fam = Ph::FAM(Ph::D1S_FN(K));
        if (fam == 0) {
            sDS[d1s_rank(K, 0) * nnp] = ps[Ph::OFF_D1S + K] * hp * val;
        } else if (fam == 1) {
            if (owned) out_g[pb[Ph::PB_GSCOL + K] + m] = ps[Ph::OFF_D1S + K] * val;
        } else {
            if (owned) red[Ph::RED_GS + d1s_rank(K, 2)] += wq * val;
        }
    }
    template <int K> __device__ __forceinline__ void H2VV(const double val) const {
        if (!WANT_H || !SECOND || !owned) return;
        if (PCX_KO_HST) { if (ps[Ph::OFF_H2VV + K] * val == 1.2345e-300) pcx_ko_sink = val; return; }
        if (regular) {
            if (MODE >= 2)
                // output order: block b of the tile's regular nodes is one contiguous
                // run of NA(b) entries per node -- the flush is a straight copy
                sH[Ph::HB_OFF(Ph::H2VV_B(K)) * PCX_THREADS + hrow * Ph::NA(Ph::H2VV_B(K))
                   + Ph::H2VV_POS(K)] = ps[Ph::OFF_H2VV + K] * val;
            else if (STG_H)
                sH[Ph::HB_OFF(Ph::H2VV_B(K)) + Ph::H2VV_POS(K)] = ps[Ph::OFF_H2VV + K] * val;
            else
                out_h[pb[Ph::PB_HREG + Ph::H2VV_B(K)] + (m - 1) * Ph::NA(Ph::H2VV_B(K))
                      + Ph::H2VV_POS(K)] = ps[Ph::OFF_H2VV + K] * val;
        } else {
            irr[K] = val;
        }
    }
    template <int K> __device__ __forceinline__ void H2VS(const double val) const {
        if (!WANT_H || !SECOND || !owned) return;
        if (regular) out_h[pb[Ph::PB_HS + K] + (m - 1)] = ps[Ph::OFF_H2VS + K] * val;
        else irr[Ph::NH2VV + K] = val;
    }
    template <int K> __device__ __forceinline__ void HTV(const double val) const {
        // rows of H against t0 / tF exist only for free times; with both times
        // fixed nothing consumes these results and the compiler drops their
        // computation (and the t-multipliers feeding it) from the body
        if (!WANT_H || !SECOND || !HAS_T || !owned) return;
        if (regular) {
            if (Ph::HAS_T0) out_h[pb[Ph::PB_HT0 + K] + (m - 1)] = ps[Ph::OFF_HT0 + K] * val;
            if (Ph::HAS_TF) out_h[pb[Ph::PB_HTF + K] + (m - 1)] = ps[Ph::OFF_HTF + K] * val;
        } else {
            irr[Ph::NH2VV + Ph::NH2VS + K] = val;
        }
    }
    template <int K> __device__ __forceinline__ void HTS(const double val) const {
        if (WANT_H && SECOND && owned) red[Ph::RED_HTS + K] += val;
    }
    template <int K> __device__ __forceinline__ void H2SS(const double val) const {
        if (WANT_H && SECOND && owned) red[Ph::RED_HSS + K] += val;
    }
};

// Small expression bodies (a few dozen results per node) fit the register file:
// there it is cheaper to let the body finish first and only then form the
// destinations, so the results are parked in registers by this sink and handed
// to the real one afterwards.
template <class Ph>
struct PcxParkingSink {
    double f[Ph::NF > 0 ? Ph::NF : 1], d1v[Ph::ND1V > 0 ? Ph::ND1V : 1];
    double d1s[Ph::ND1S > 0 ? Ph::ND1S : 1], h2vv[Ph::NH2VV > 0 ? Ph::NH2VV : 1];
    double h2vs[Ph::NH2VS > 0 ? Ph::NH2VS : 1], h2ss[Ph::NH2SS > 0 ? Ph::NH2SS : 1];
    double htv[Ph::NHTV > 0 ? Ph::NHTV : 1], hts[Ph::NHTS > 0 ? Ph::NHTS : 1];
    template <int I> __device__ __forceinline__ void F(const double v) { f[I] = v; }
    template <int K> __device__ __forceinline__ void D1V(const double v) { d1v[K] = v; }
    template <int K> __device__ __forceinline__ void D1S(const double v) { d1s[K] = v; }
    template <int K> __device__ __forceinline__ void H2VV(const double v) { h2vv[K] = v; }
    template <int K> __device__ __forceinline__ void H2VS(const double v) { h2vs[K] = v; }
    template <int K> __device__ __forceinline__ void H2SS(const double v) { h2ss[K] = v; }
    template <int K> __device__ __forceinline__ void HTV(const double v) { htv[K] = v; }
    template <int K> __device__ __forceinline__ void HTS(const double v) { hts[K] = v; }
};

template <int... Is> struct PcxSeq {};
template <int N, int... Is> struct PcxMakeSeq : PcxMakeSeq<N - 1, N - 1, Is...> {};
template <int... Is> struct PcxMakeSeq<0, Is...> { typedef PcxSeq<Is...> type; };

#define PCX_DRAIN(NAME, FIELD)                                                            \
    template <class Ph, int... Is>                                                        \
    __device__ __forceinline__ void pcx_drain_##NAME(const PcxParkingSink<Ph>& a,         \
                                                     const PcxNodeSink<Ph>& o, PcxSeq<Is...>) { \
        int dummy[] = {0, (o.template NAME<Is>(a.FIELD[Is]), 0)...};                      \
        (void)dummy;                                                                      \
    }
PCX_DRAIN(F, f) PCX_DRAIN(D1V, d1v) PCX_DRAIN(D1S, d1s) PCX_DRAIN(H2VV, h2vv)
PCX_DRAIN(H2VS, h2vs) PCX_DRAIN(H2SS, h2ss) PCX_DRAIN(HTV, htv) PCX_DRAIN(HTS, hts)
#undef PCX_DRAIN

// ---------------------------------------------------------------------------
// One tile of phase Ph.  A tile that writes something the border pass reads
// (reduction partials, end-node values) or overwrites (gradient zeros) fences
// those writes and bumps the instance's ticket as soon as its node phase ends.
// ---------------------------------------------------------------------------
// The thread's first scatter work item (decoded in the prologue) and the G
// constants: ONE static block shared by every phase instantiation of pcx_tile
// (declared inside the template these arrays were allocated once per phase: 20 KB of
// static shared memory for a 4-phase problem, which cost a resident CTA per SM).
struct PcxTileStatic {
    double coef[2 * PCX_THREADS];
    i64 o[PCX_THREADS];
    int dp[PCX_THREADS], dstep[PCX_THREADS], ostep[PCX_THREADS], cnt[PCX_THREADS];
    double cst[1 + 2 * PCX_NY_MAX];
    // tables of the run being scattered (period offsets per variable, slot bases,
    // section range): staged once per run so that decoding a work item needs ONE
    // global load (its recipe word) instead of a chain of four
    i64 run_gb[PCX_NV_MAX + 1];
    int run_tv[PCX_NV_MAX + 2];
    int run_lohi[2];
};

template <class Ph>
__device__ __forceinline__ void pcx_stage_run(const PcxParams& p, const int run, const int tid,
                                              const unsigned long long keep, PcxTileStatic& ts) {
    const int* tv = p.type_var_off + pcx_ld_keep(p.run_type + run, keep) * (p.nvmax + 1);
    for (int i = tid; i <= Ph::NV; i += PCX_THREADS) {
        ts.run_tv[i] = tv[i];
        if (i < Ph::NV) ts.run_gb[i] = pcx_ld_keep(p.run_gbase + (i64)run * p.nvmax + i, keep);
    }
    if (tid == 0) {
        ts.run_lohi[0] = pcx_ld_keep(p.run_slo + run, keep);
        ts.run_lohi[1] = pcx_ld_keep(p.run_shi + run, keep);
    }
}

// run setup / slot decoding from the staged tables (shared memory only, apart
// from the recipe word the caller has already fetched)
template <class Ph>
__device__ __forceinline__ bool pcx_run_setup_s(const PcxTileStatic& ts, const int tid,
                                                const int* sSecOrder, PcxRun& R) {
    constexpr int T = PCX_THREADS;
    R.s_lo = ts.run_lohi[0];
    R.s_hi = ts.run_lohi[1];
    R.tv = nullptr;
    R.rec0 = ts.run_tv[0];
    R.Ptot = ts.run_tv[Ph::NV] - R.rec0;
    if (R.Ptot == 0) return false;
    R.per = 1; R.u0 = tid; R.sc0 = R.s_lo;
    if (R.Ptot < T) {
        R.per = T / R.Ptot;
        const int q = tid / R.Ptot;
        R.u0 = tid - q * R.Ptot;
        R.sc0 = R.s_lo + q;
        if (q >= R.per) return false;
    }
    R.nstep = sSecOrder[R.s_lo + 1] - 1;
    return true;
}

template <class Ph>
__device__ __forceinline__ bool pcx_decode_word(const unsigned long long w, const PcxRun& R,
                                                const PcxTileStatic& ts, const double* sB,
                                                const int* sSecNode, const int sD_off,
                                                const int sDP_off, const int nnp, const int nsp,
                                                PcxSlot& S) {
    const u32 lo = (u32)w;
    if (lo >> RC_SKIP_BIT) return false;
    const int a = (int)((w >> RC_VAR_SHIFT) & 0xffu);
    const int local = (int)((w >> RC_LOCAL_SHIFT) & ((1u << RC_LOCAL_BITS) - 1));
    const int Pa = ts.run_tv[a + 1] - ts.run_tv[a];
    const int e = lo & ((1u << RC_E_BITS) - 1);
    const int bi = (lo >> RC_B_SHIFT) & ((1u << RC_B_BITS) - 1);
    const int mloc = (int)((w >> RC_M_SHIFT) & ((1u << RC_M_BITS) - 1));
    S.cc = ts.cst[(lo >> RC_C_SHIFT) & ((1u << RC_C_BITS) - 1)];
    const bool prev = (lo >> RC_PREV_BIT) & 1u;
    S.bcoef = sB[bi];
    S.dp = 0; S.dstep = 0;
    if (e) {
        S.dp = prev ? sDP_off + (e - 1) * nsp + R.sc0
                    : sD_off + (e - 1) * nnp + mloc + sSecNode[R.sc0 + 1];
        S.dstep = prev ? R.per : R.per * R.nstep;
    }
    S.o = ts.run_gb[a] + local + (i64)(R.sc0 - R.s_lo) * Pa;
    S.ostep = R.per * Pa;
    S.cnt = (R.s_hi - R.sc0 + R.per - 1) / R.per;
    return true;
}

// Staged node-diagonal Hessian entries -> global memory, one variable block at a
// time: the block's slots of the tile's regular nodes are contiguous
// (NA(b) per node), consecutive threads write consecutive doubles.
template <class Ph, int HSTR, int... Bs>
__device__ __forceinline__ void pcx_flush_h(const double* sH, double* out_h, const i64* pb,
                                            const i64 m_first, const int a0, const int n_reg,
                                            const int tid, PcxSeq<Bs...>) {
    constexpr int T = PCX_THREADS;
    int dummy[] = {0, ([&] {
        constexpr int NAb = Ph::NA(Bs);
        if (NAb > 0) {
            double* dst = out_h + pb[Ph::PB_HREG + Bs] + (m_first - 1) * NAb;
            const double* src = sH + a0 * HSTR + Ph::HB_OFF(Bs);
            for (int i = tid; i < n_reg * NAb; i += T) {
                const int nd = i / NAb;
                dst[i] = src[nd * HSTR + (i - nd * NAb)];
            }
        }
    }(), 0)...};
    (void)dummy;
}

// two-pass form: the entries were staged in output order, block b at sH + HB_OFF(b) * T
template <class Ph, int... Bs>
__device__ __forceinline__ void pcx_copy_h(const double* sH, double* out_h, const i64* pb,
                                           const i64 m_first, const int n_reg, const int tid,
                                           PcxSeq<Bs...>) {
    constexpr int T = PCX_THREADS;
    int dummy[] = {0, ([&] {
        constexpr int NAb = Ph::NA(Bs);
        if (NAb > 0) {
            // trip counts are 1..NA(b): no unrolling (the unrolled form spent 20 % of the
            // kernel's instructions on prologues / remainders of ten short loops)
            double* dst = out_h + pb[Ph::PB_HREG + Bs] + (m_first - 1) * NAb + tid;
            const double* src = sH + Ph::HB_OFF(Bs) * T + tid;
            const int n = n_reg * NAb - tid;
#pragma unroll 1
            for (int i = 0; i < n; i += T) dst[i] = src[i];
        }
    }(), 0)...};
    (void)dummy;
}

template <class Ph>
__device__ void pcx_tile(const PcxParams& p, const int tile, const int inst, const int phase,
                         unsigned char* smem_raw, PcxTileStatic& ts)
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
    constexpr int NOUT = NF + Ph::ND1V + Ph::ND1S + Ph::NH2VV + Ph::NH2VS + Ph::NH2SS
                         + Ph::NHTV + Ph::NHTS;
    constexpr bool PARK = NOUT <= 40;              // small bodies: results parked in registers
    constexpr bool TWO = (PCX_TWO_PASS != 0) && WANT_G && WANT_H && !PARK
                         && Ph::NH2VV >= PCX_TWO_PASS_MIN && !Ph::STAGE_H;
    // launches without the Jacobian (the Hessian callback of a host solver) stage nothing
    // for a scatter: the same full-sector Hessian flush fits next to the multipliers in
    // a single pass
    constexpr bool HST = (PCX_TWO_PASS != 0) && !WANT_G && WANT_H && !PARK
                         && Ph::NH2VV >= PCX_TWO_PASS_MIN && !Ph::STAGE_H;

    const int tid = threadIdx.x;
    PCX_STAMP(0);
    // An instantiation that serves several phases (identical bodies up to a few literals,
    // codegen.py: share_groups) finds the tables of the phase at hand at run time; one that
    // serves a single phase keeps compile-time constant-bank addresses.
    const i64* pb = pcx_c_pbase + (Ph::SHARED ? PCX_PHASE_PBASE(phase) : Ph::PBASE_OFF);
    const double* ps = pcx_c_pscal + (Ph::SHARED ? PCX_PHASE_PSCAL(phase) : Ph::PSCAL_OFF);
    const double* kc = ps + Ph::OFF_KC;
    const i64 N = pb[Ph::PB_N], K = pb[Ph::PB_K];
    const i64 xo = pb[Ph::PB_XOFF], co = pb[Ph::PB_COFF];
    const i64 sec_off = pb[Ph::PB_SECOFF];         // offset into sec_order/h/type
    const i64* sec_node = p.sec_node + sec_off + (Ph::SHARED ? phase : Ph::INDEX);   // K+1 per phase
    const unsigned long long keep = pcx_policy_keep();
    const double* x = p.x + (i64)inst * p.num_x;
    const double* lam = WANT_H ? p.lam + (i64)inst * p.num_c : nullptr;
    {
        // Speculative L2 prefetch of the tile's iterate/multiplier lines, issued
        // before the tile descriptor (an L2 round trip) is known: tiles are
        // balanced by node count (structure.py:_balanced_edges), so the first
        // node of tile j of a phase is within one section of j*(N-1)/ntiles.
        // A wrong guess costs nothing; a right one turns the DRAM latency of
        // the dependent loads below into an L2 hit.
        const int t_lo = (int)pb[Ph::PB_TILE0], nt = (int)pb[Ph::PB_TILE1] - t_lo;
        const i64 est = (i64)((float)(tile - t_lo) * ((float)(N - 1) / (float)nt));
        constexpr int NL = (T + 32) / 16;              // 128 B lines per stream
        constexpr int NSTREAM = NV + (WANT_H ? NY + NP : 0);
        for (int i = tid; i < NSTREAM * NL; i += T) {
            const int st = i / NL;
            i64 off = est + (i64)(i - st * NL) * 16;
            if (off > N - 2) off = N - 2;
            if (off < 0) off = 0;
            const double* ptr;
            if (st < NV) ptr = x + xo + (i64)st * N + off;
            else if (st < NV + NY) ptr = lam + co + (i64)(st - NV) * (N - 1) + off;
            else ptr = lam + co + (i64)NY * (N - 1) + (i64)(st - NV - NY) * N + off;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ptr));
        }
    }
    const i64* td = p.tile_desc + (i64)tile * 8;   // phase,k0,k1,node0,nn,run0,run1,prev_rows
    const int k0 = (int)pcx_ld_keep(td + 1, keep), k1 = (int)pcx_ld_keep(td + 2, keep);
    const int nsec = k1 - k0;
    const i64 node0 = pcx_ld_keep(td + 3, keep);
    const int nn = (int)pcx_ld_keep(td + 4, keep);
    const int run0 = (int)pcx_ld_keep(td + 5, keep);
    const int nruns = (int)pcx_ld_keep(td + 6, keep) - run0;
    const int prev_rows = (int)pcx_ld_keep(td + 7, keep);   // defect rows of section k0-1
    const bool last_tile = (k1 == (int)K);
    const bool has_prev = (k0 > 0);
    PCX_STAMP(7);
#ifdef PCX_DEBUG_TIMELINE
    pcx_stamp_smid(p, tile);
#endif
    const int nnp = nn | 1;                        // odd stride: no bank conflicts
    // Multi-wave grids (PCX_EARLY_WAIT, chosen by the host when the tiles need three
    // waves or more): only the first wave can overlap a previous kernel, so the
    // dependency wait moves up here and the thread's iterate / multiplier loads are
    // put in flight BEFORE the table chain below instead of after it -- their DRAM
    // latency hides behind the four L2 round trips of the tables.
    double xt0[NV > 0 ? NV : 1];
    double lam_p[NP > 0 ? NP : 1], lam_q[NQ > 0 ? NQ : 1];     // path (this node) / integral rows
    double xt_t0 = 0.0, xt_tF = 0.0;
    if (PCX_EARLY_WAIT) {
        if (!p.independent) pcx_grid_dependency_wait();
        pcx_first_wave_stagger();
#pragma unroll
        for (int a = 0; a < NV; ++a)
            xt0[a] = (tid < nn) ? pcx_ld_f64(x + xo + (i64)a * N + node0 + tid) : 0.0;
        if (Ph::HAS_T0) xt_t0 = pcx_ld_f64(x + pb[Ph::PB_T0X]);
        if (Ph::HAS_TF) xt_tF = pcx_ld_f64(x + pb[Ph::PB_TFX]);
        if (WANT_H) {
#pragma unroll
            for (int j = 0; j < NP; ++j)
                lam_p[j] = (tid < nn) ? pcx_ld_f64(lam + co + (i64)NY * (N - 1) + (i64)j * N + node0 + tid) : 0.0;
#pragma unroll
            for (int i = 0; i < NQ; ++i) lam_q[i] = pcx_ld_f64(lam + co + (i64)NY * (N - 1) + (i64)NP * N + i);
        }
    }

    // ---- shared memory carve-up ------------------------------------------
    double* sB = reinterpret_cast<double*>(smem_raw);          // btab
    double* sHk = sB + p.btab_len;                             // nsec+1 (prev first)
    double* sF = sHk + (nsec + 1);                             // NY * nnp
    double* sD = sF + (NEED_SF ? NY * nnp : 0);                // ND1V * nnp
    double* sDP = sD + (WANT_G ? Ph::ND1V * nnp : 0);          // ND1V * (nsec + 1)
    const int nsp = nsec + 1;
    double* sDS = sDP + (WANT_G ? Ph::ND1V * nsp : 0);         // NDS * nnp
    double* sLam = sDS + (WANT_G ? NDS * nnp : 0);             // NY * (nn + PCX_LAM_HALO)
    const int lam_stride = nn + PCX_LAM_HALO;
    double* sH = sLam + (WANT_H ? NY * lam_stride : 0);        // HP * nn (staged H entries)
    double* sRed = sH + ((WANT_H && Ph::STAGE_H) ? Ph::HP * nn : 0);   // T/32
    // two-pass node phase: the second pass stages NH2VV * T Hessian entries over
    // sD / sDP / sDS / sLam, all dead once the Jacobian has been scattered
    if (TWO && sD + Ph::NH2VV * T > sRed) sRed = sD + Ph::NH2VV * T;
    if (HST) sRed = sH + Ph::NH2VV * T;
    int* sSecNode = reinterpret_cast<int*>(sRed + T / 32);     // nsec+2 (prev first)
    int* sSecOrder = sSecNode + (nsec + 2);                    // nsec+1 (prev first)
    int* sNodeSec = sSecOrder + (nsec + 1);                    // nn
    int* sAoff = sNodeSec + nn;                                // nsec+1 (prev first): A(N_k) in sB
    int* sWoff = sAoff + (nsec + 1);                           // nsec+1 (prev first): w(N_k) in sB
    // the thread's first scatter work item, decoded in the prologue
    double* sPreCoef = ts.coef;
    i64* sPreO = ts.o;
    int *sPreDp = ts.dp, *sPreDstep = ts.dstep, *sPreOstep = ts.ostep, *sPreCnt = ts.cnt;
    double* sCst = ts.cst;

    // ---- table-only part of the prologue (before the dependency wait) --------
    // quadrature table, G constants and the section table slice of the tile go to
    // shared memory, the node -> section map is built: none of it depends on the
    // iterate, so under programmatic dependent launch all of it overlaps the
    // previous kernel's drain
    for (int i = tid; i < p.btab_len; i += T) sB[i] = pcx_ld_keep(p.btab + i, keep);
    if (WANT_G && tid < 1 + 2 * NY) sCst[tid] = ps[Ph::OFF_GCST + tid];
    for (int s = tid; s <= nsec; s += T) {
        const int k = k0 - 1 + s;                  // s = 0 is the previous section
        const bool ok = (k >= 0);
        sHk[s] = ok ? pcx_ld_keep(p.sec_h + sec_off + k, keep) : 0.0;
        const int ord = ok ? pcx_ld_keep(p.sec_order + sec_off + k, keep) : 0;
        sSecOrder[s] = ord;
        sSecNode[s] = ok ? (int)(pcx_ld_keep(sec_node + k, keep) - node0) : 0;
        // where the section's quadrature tables start in sB: looked up here so the
        // node phase reads shared memory only
        sAoff[s] = p.order_a_off[ord];
        sWoff[s] = p.order_w_off[ord];
    }
    if (tid == 0) sSecNode[nsec + 1] = nn - 1;
    if (WANT_G && PCX_PRE_STAGED && nruns > 0) pcx_stage_run<Ph>(p, run0, tid, keep, ts);
    __syncthreads();
    for (int s = tid; s < nsec; s += T) {
        const int b = sSecNode[s + 1], n = sSecOrder[s + 1];
        for (int m = 0; m < n - 1; ++m) sNodeSec[b + m] = s;
        if (s == nsec - 1) sNodeSec[b + n - 1] = s;
    }
    __syncthreads();
    // ---- this thread's node, as far as the tables know it (a tile never holds
    // more nodes than the CTA has threads): its section, position, section
    // lengths and quadrature weight are all formed before the dependency wait
    const int ml = tid;
    const bool active = ml < nn;
    const int s = active ? sNodeSec[ml] : 0;
    const int mloc = ml - sSecNode[s + 1];
    const int n_k = sSecOrder[s + 1];
    const double h_k = sHk[s + 1];
    const bool start_with_prev = active && (mloc == 0) && (s > 0 || has_prev);
    const int n_pr = start_with_prev ? sSecOrder[s] : 0;
    const double h_pr = start_with_prev ? sHk[s] : 0.0;
    const int a_k = sAoff[s + 1] + mloc, a_pr = sAoff[s] + n_pr - 1;
    const int lrow_k = prev_rows + sSecNode[s + 1], lrow_pr = prev_rows + sSecNode[s];
    // quadrature weight of this node, accumulated as pycollo/mesh.py:325-326
    double wq = 0.0;
    if (NQ > 0 && (WANT_C || WANT_G || WANT_H)) {
        if (start_with_prev)
            wq = __dmul_rn(sB[sWoff[s] + n_pr - 1], h_pr);
        wq = __dadd_rn(wq, __dmul_rn(sB[sWoff[s + 1] + mloc], h_k));
    }
    // the scatter's tables (run descriptor -> type -> recipe word -> slot base) are
    // a chain of dependent loads: walked here, in the shadow of the previous
    // kernel, instead of between the node phase and the first store
    bool pre_more = false;
    if (WANT_G && PCX_PRE_STAGED) {
        // the first run's tables were staged with the section table (before the first
        // barrier): the decode is ONE global load (the recipe word) + shared memory
        sPreCnt[tid] = 0;
        if (nruns > 0) {
            PcxRun run;
            if (pcx_run_setup_s<Ph>(ts, tid, sSecOrder, run)) {
                const unsigned long long w = pcx_ld_keep(p.recipes + run.rec0 + run.u0, keep);
                PcxSlot sl;
                if (pcx_decode_word<Ph>(w, run, ts, sB, sSecNode, (int)(sD - sB), (int)(sDP - sB),
                                        nnp, nsp, sl)) {
                    sPreCoef[2 * tid] = sl.bcoef; sPreCoef[2 * tid + 1] = sl.cc;
                    sPreO[tid] = sl.o; sPreDp[tid] = sl.dp; sPreDstep[tid] = sl.dstep;
                    sPreOstep[tid] = sl.ostep; sPreCnt[tid] = sl.cnt;
                }
            }
            pre_more = (nruns > 1) || (ts.run_tv[NV] - ts.run_tv[0] > T);
        }
    } else if (WANT_G) {
        sPreCnt[tid] = 0;
        if (nruns > 0) {
            PcxRun run;
            if (pcx_run_setup<Ph>(p, run0, tid, keep, sSecOrder, run)) {
                PcxSlot sl;
                if (pcx_decode_slot<Ph>(p, run0, run.u0, run, keep, sB, sCst, sSecNode,
                                        (int)(sD - sB), (int)(sDP - sB), nnp, nsp, sl)) {
                    sPreCoef[2 * tid] = sl.bcoef; sPreCoef[2 * tid + 1] = sl.cc;
                    sPreO[tid] = sl.o; sPreDp[tid] = sl.dp; sPreDstep[tid] = sl.dstep;
                    sPreOstep[tid] = sl.ostep; sPreCnt[tid] = sl.cnt;
                }
            }
            // CTA-uniform: is there anything beyond one work item per thread?
            const int* tv0 = p.type_var_off + pcx_ld_keep(p.run_type + run0, keep) * (p.nvmax + 1);
            pre_more = (nruns > 1) || (tv0[NV] - tv0[0] > T);
            if (pre_more && !PCX_DECODE_V1) pcx_stage_run<Ph>(p, run0, tid, keep, ts);   // visible after the next barrier
        }
    }
    // ---- the iterate and the multipliers ------------------------------------------
    // everything above touched only the engine's immutable tables; x, lam and
    // every output may belong to the previous kernel in the stream, which has to
    // be complete from here on -- unless the caller declared the evaluations
    // independent (distinct buffers, rotating scratch): then consecutive kernels
    // overlap freely and their load / compute / store phases interleave
    if (!PCX_EARLY_WAIT) {
        if (!p.independent) pcx_grid_dependency_wait();
        pcx_first_wave_stagger();
        PCX_STAMP(8);
#pragma unroll
        for (int a = 0; a < NV; ++a)
            xt0[a] = (tid < nn) ? x[xo + (i64)a * N + node0 + tid] : 0.0;
        if (Ph::HAS_T0) xt_t0 = x[pb[Ph::PB_T0X]];
        if (Ph::HAS_TF) xt_tF = x[pb[Ph::PB_TFX]];
        if (WANT_H) {
#pragma unroll
            for (int j = 0; j < NP; ++j)
                lam_p[j] = (tid < nn) ? lam[co + (i64)NY * (N - 1) + (i64)j * N + node0 + tid] : 0.0;
#pragma unroll
            for (int i = 0; i < NQ; ++i) lam_q[i] = lam[co + (i64)NY * (N - 1) + (i64)NP * N + i];
        }
    }
    if (WANT_H) {
        // multipliers of the defect rows of sections k0-1 .. k1-1
        const int nrows = nn - 1 + prev_rows;
        const double* lp = lam + co + (node0 - prev_rows) + tid;
#pragma unroll 1
        for (int r = tid; r < nrows; r += T, lp += T) {
#pragma unroll
            for (int i = 0; i < NY; ++i) sLam[i * lam_stride + r] = lp[(i64)i * (N - 1)];
        }
    }
    __syncthreads();
    PCX_STAMP(1);

    // ---- phase scalars ------------------------------------------------------
    const double t0 = pcx_unscale(ps[Ph::OFF_TINFO + 0], xt_t0, ps[Ph::OFF_TINFO + 1]);
    const double tF = pcx_unscale(ps[Ph::OFF_TINFO + 2], xt_tF, ps[Ph::OFF_TINFO + 3]);
    const double hp = 0.5 * (tF - t0);
    double sv[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < NS; ++j)
        sv[j] = pcx_unscale(pcx_c_gscal[PCX_GS_VS + j], x[p.num_x - NS + j], pcx_c_gscal[PCX_GS_RS + j]);

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
    // (declared at function scope: the two-pass form evaluates the body a second time
    // after the Jacobian scatter, from the same unscaled variables and multipliers)
    double v[NV + NS > 0 ? NV + NS : 1];
    double muh[NF > 0 ? NF : 1], mut[NF > 0 ? NF : 1];
    const bool owned = active && ((ml < nn - 1) || last_tile);
    const i64 m = node0 + ml;
#define PCX_SINK_SETUP(SK)                                                                   \
    SK.ps = ps; SK.pb = pb;                                                                   \
    SK.sF = sF + ml; SK.sD = sD + ml; SK.sDS = sDS + ml; SK.sDP = sDP + s;                    \
    SK.sH = sH + ml * Ph::HP;                                                                 \
    SK.nnp = nnp; SK.nsp = nsp; SK.sec_start = (mloc == 0);                                   \
    SK.hp = hp; SK.wq = wq; SK.h_k = h_k; SK.h_pr = h_pr;                                     \
    SK.owned = owned; SK.regular = (m != 0) && (m != N - 1);                                  \
    SK.m = m; SK.N = N;                                                                       \
    SK.out_c_path = WANT_C ? out_c + co + (i64)NY * (N - 1) + m : nullptr;                    \
    SK.out_dy = WANT_DY ? out_dy + pb[Ph::PB_DYOFF] + m : nullptr;                            \
    SK.out_g = out_g; SK.out_h = out_h;                                                       \
    SK.irr = bv + pb[m == 0 ? Ph::PB_IRR0 : Ph::PB_IRR1];                                     \
    SK.red = red;
    if (active) {
#pragma unroll
        for (int a = 0; a < NV; ++a)
            v[a] = pcx_unscale(ps[Ph::OFF_VV + a], xt0[a], ps[Ph::OFF_RV + a]);
#pragma unroll
        for (int j = 0; j < NS; ++j) v[NV + j] = sv[j];

#pragma unroll
        for (int e = 0; e < NF; ++e) { muh[e] = 0.0; mut[e] = 0.0; }
        if (WANT_H) {
            double lacc[NY > 0 ? NY : 1];
#pragma unroll
            for (int i = 0; i < NY; ++i) lacc[i] = 0.0;
            if (start_with_prev) {
                const double* Apr = sB + a_pr;
                const double* lrow = sLam + lrow_pr;
                for (int l = 0; l < n_pr - 1; ++l) {
                    const double cl = __dmul_rn(Apr[l * n_pr], h_pr);
#pragma unroll
                    for (int i = 0; i < NY; ++i) lacc[i] += lrow[i * lam_stride + l] * cl;
                }
            }
            if (owned) {
                const double* Ak = sB + a_k;
                const double* lrow = sLam + lrow_k;
                for (int l = 0; l < n_k - 1; ++l) {
                    const double cl = __dmul_rn(Ak[l * n_k], h_k);
#pragma unroll
                    for (int i = 0; i < NY; ++i) lacc[i] += lrow[i * lam_stride + l] * cl;
                }
            }
#pragma unroll
            for (int i = 0; i < NY; ++i) {
                const double mu = ps[Ph::OFF_WFN + i] * lacc[i];
                mut[i] = mu;
                muh[i] = hp * mu;
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const double mu = owned ? ps[Ph::OFF_WFN + NY + j]
                    * lam_p[j] : 0.0;
                muh[NY + j] = mu;
                mut[NY + j] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const double mu = -ps[Ph::OFF_WFN + NY + NP + i] * lam_q[i] * wq;
                mut[NY + NP + i] = mu;
                muh[NY + NP + i] = hp * mu;
            }
        }

        // the generated body hands every result to the sink the moment it
        // exists: staged (G), stored (H, dy, path rows of c) or accumulated
        if (PARK) {
            PcxParkingSink<Ph> parked;
            Ph::eval(v, muh, mut, kc, parked);
            PcxNodeSink<Ph> sink;
            PCX_SINK_SETUP(sink)
            pcx_drain_F(parked, sink, typename PcxMakeSeq<NF>::type());
            pcx_drain_D1V(parked, sink, typename PcxMakeSeq<Ph::ND1V>::type());
            pcx_drain_D1S(parked, sink, typename PcxMakeSeq<Ph::ND1S>::type());
            pcx_drain_H2VV(parked, sink, typename PcxMakeSeq<Ph::NH2VV>::type());
            pcx_drain_H2VS(parked, sink, typename PcxMakeSeq<Ph::NH2VS>::type());
            pcx_drain_H2SS(parked, sink, typename PcxMakeSeq<Ph::NH2SS>::type());
            pcx_drain_HTV(parked, sink, typename PcxMakeSeq<Ph::NHTV>::type());
            pcx_drain_HTS(parked, sink, typename PcxMakeSeq<Ph::NHTS>::type());
        } else if (TWO) {
            PcxNodeSink<Ph, 1> sink;                 // first pass: f, first derivatives
            PCX_SINK_SETUP(sink)
            Ph::eval(v, muh, mut, kc, sink);
        } else if (HST) {
            PcxNodeSink<Ph, 3> sink;                 // single pass, Hessian staged for the flush
            PCX_SINK_SETUP(sink)
            sink.sH = sH;
            sink.hrow = ml - ((node0 == 0) ? 1 : 0);
            Ph::eval(v, muh, mut, kc, sink);
        } else {
            PcxNodeSink<Ph> sink;
            PCX_SINK_SETUP(sink)
            Ph::eval(v, muh, mut, kc, sink);
        }
        if (WANT_GRAD && owned) {
#pragma unroll
            for (int a = 0; a < NV; ++a) out_grad[xo + (i64)a * N + m] = 0.0;
        }
    }
#ifdef PCX_DEBUG_TIMELINE
    if ((tid & 31) == 0 && (tid >> 5) < 4) {           // per-warp end of the node phase
        unsigned long long tw;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tw));
        p.partials[(i64)tile * 16 + 11 + (tid >> 5)] = (double)(tw & ((1ull << 40) - 1));
    }
#endif
    if (WANT_GRAD && tile == 0) {
        // non-node entries (q, t of every phase, s) are zeroed once; the border
        // pass overwrites the structural ones afterwards
        for (int i = tid; i < PCX_NPOINT; i += T) out_grad[p.pt_x[i]] = 0.0;
    }
    __syncthreads();

    PCX_STAMP(2);
    // ---- reductions -> per-tile partials, and the signal to the border CTA ------
    // Everything the border pass reads from a tile (partials, end-node values,
    // zeroed gradient entries) exists once the node phase is over: signalling
    // here, before the bulk of the value stores is issued, keeps the fence
    // (which waits for the CTA's outstanding stores) short.
    auto signal_border = [&]() {
        if (pcx_need_red<Ph>()) {
    #pragma unroll
            for (int k = 0; k < Ph::NRED; ++k) {
                const double r = pcx_block_sum(red[k], sRed);
                if (tid == 0)
                    p.partials[((i64)inst * p.num_tiles + tile) * p.nred_max + k] = r;
            }
        }
        if (pcx_need_red<Ph>() || (WANT_H && (k0 == 0 || last_tile))
            || (WANT_GRAD && (k0 == 0 || last_tile || tile == 0))) {
            // only the threads that wrote something the border pass reads fence: thread 0
            // (reduction partials), the threads of the phase's two end nodes (their
            // Hessian entries) and, with GRAD, every thread (zeroed gradient entries).
            // A fence waits for the calling thread's outstanding stores -- for a large
            // body that is ~40 direct Hessian stores per node, and a CTA-wide fence made
            // the signalling tiles twice as slow as the others.
            const i64 m_me = node0 + tid;
            if (WANT_GRAD || tid == 0 || (active && (m_me == 0 || m_me == N - 1))) __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(p.ticket + inst, 1u);
        }
    };
    if (!TWO) signal_border();
    if (HST) {
        const int a0 = (node0 == 0) ? 1 : 0;               // node 0 / N-1 go through irr
        pcx_copy_h<Ph>(sH, out_h, pb, node0 + a0, (nn - 1) - a0, tid,
                       typename PcxMakeSeq<NV>::type());
    }
    // ---- staged Hessian entries of the tile's regular nodes, block by block ------
    if (WANT_H && Ph::STAGE_H) {
        const int a0 = (node0 == 0) ? 1 : 0;               // node 0 / N-1 go through irr
        pcx_flush_h<Ph, Ph::HP>(sH, out_h, pb, node0 + a0, a0, (nn - 1) - a0, tid,
                                typename PcxMakeSeq<NV>::type());
    }
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
            const double* Arow = sB + sAoff[s + 1] + l * n_k;
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

    // ---- coalesced scatter of the Jacobian values (see pcx_store_run) ----------
    unsigned token = 0;
    if (TWO && PCX_STORE_TOKEN) token = pcx_store_token_acquire(tid);
    if (WANT_G) {
#if PCX_DECODE_V1
        // (round-1 form, kept for A/B: every work item walks the table chain)
        if (sPreCnt[tid] > 0)
            pcx_store_run(out_g + sPreO[tid], sPreOstep[tid], sB + sPreDp[tid], sPreDstep[tid],
                          sPreCoef[2 * tid], sPreCoef[2 * tid + 1], sPreCnt[tid]);
        if (pre_more) {
#pragma unroll 1
            for (int r = 0; r < nruns; ++r) {
                PcxRun run;
                if (!pcx_run_setup<Ph>(p, run0 + r, tid, keep, sSecOrder, run)) continue;
                for (int u = run.u0 + (r == 0 ? T : 0); u < run.Ptot; u += T) {
                    PcxSlot sl;
                    if (pcx_decode_slot<Ph>(p, run0 + r, u, run, keep, sB, sCst, sSecNode,
                                            (int)(sD - sB), (int)(sDP - sB), nnp, nsp, sl))
                        pcx_store_run(out_g + sl.o, sl.ostep, sB + sl.dp, sl.dstep, sl.bcoef, sl.cc, sl.cnt);
                    if (run.Ptot < T) break;
                }
            }
        }
    }
#else
        if (!pre_more) {
            // the thread's only work item was decoded in the prologue
            if (sPreCnt[tid] > 0)
                pcx_store_run(out_g + sPreO[tid], sPreOstep[tid], sB + sPreDp[tid], sPreDstep[tid],
                              sPreCoef[2 * tid], sPreCoef[2 * tid + 1], sPreCnt[tid]);
        } else {
            // several work items per thread (large bodies, ragged meshes): the run's
            // tables are staged in shared memory and the recipe word of the NEXT item
            // is fetched before the current item's stores are issued, so the only
            // global load of a decode is never waited for
#pragma unroll 1
            for (int r = 0; r < nruns; ++r) {
                if (r > 0) {
                    __syncthreads();
                    pcx_stage_run<Ph>(p, run0 + r, tid, keep, ts);
                    __syncthreads();
                }
                PcxRun run;
                const bool ok = pcx_run_setup_s<Ph>(ts, tid, sSecOrder, run);
                int u = run.u0 + (r == 0 ? T : 0);       // run0's first item: pre-decoded
                bool have = ok && u < run.Ptot;
                unsigned long long w = have ? pcx_ld_keep(p.recipes + run.rec0 + u, keep) : 0ull;
                if (r == 0 && sPreCnt[tid] > 0)
                    pcx_store_run(out_g + sPreO[tid], sPreOstep[tid], sB + sPreDp[tid],
                                  sPreDstep[tid], sPreCoef[2 * tid], sPreCoef[2 * tid + 1],
                                  sPreCnt[tid]);
                while (have) {
                    const int un = u + T;
                    const bool more = (run.Ptot >= T) && (un < run.Ptot);
                    const unsigned long long wn =
                        more ? pcx_ld_keep(p.recipes + run.rec0 + un, keep) : 0ull;
                    PcxSlot sl;
                    if (pcx_decode_word<Ph>(w, run, ts, sB, sSecNode, (int)(sD - sB),
                                            (int)(sDP - sB), nnp, nsp, sl))
                        pcx_store_run(out_g + sl.o, sl.ostep, sB + sl.dp, sl.dstep, sl.bcoef,
                                      sl.cc, sl.cnt);
                    w = wn; u = un; have = more;
                }
            }
        }
    }
#endif

    // ---- second pass of the two-pass node phase ----------------------------------
    if (TWO) {
        __syncthreads();                 // the staged first derivatives have been consumed
        if (PCX_STORE_TOKEN) pcx_store_token_release(tid, token);
        PCX_STAMP(4);
        if (active) {
            PcxNodeSink<Ph, 2> sink;     // second derivatives only; the compiler drops the rest
            PCX_SINK_SETUP(sink)
            sink.sH = sD;
            sink.hrow = ml - ((node0 == 0) ? 1 : 0);
            Ph::eval(v, muh, mut, kc, sink);
        }
        __syncthreads();
        PCX_STAMP(15);
        signal_border();
        const int a0 = (node0 == 0) ? 1 : 0;               // node 0 / N-1 go through irr
        pcx_copy_h<Ph>(sD, out_h, pb, node0 + a0, (nn - 1) - a0, tid,
                       typename PcxMakeSeq<NV>::type());
    }
#undef PCX_SINK_SETUP

    PCX_STAMP(3);
}

// ---------------------------------------------------------------------------
// Border pass: one dedicated CTA per instance (blockIdx.x == num_tiles, the
// last one dispatched).  Everything that does not depend on the tiles -- point
// variables, the point function, the multipliers, the border-map tables -- is
// evaluated while the tiles run; the CTA then waits on the instance's ticket
// until every tile that feeds it (reduction partials, end-node values, zeroed
// gradient entries) has signalled, and applies the border map.  When no
// reduction is selected only the first and last tile of each phase signal, so
// the pass is off the kernel's critical path.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pcx_ld_acquire_sys(const unsigned long long* ptr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pcx_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// a peer that never arrives (failed launch on another rank) must not hang this
// GPU: the spin gives up after 5 s and leaves 1 in the engine's status word
#define PCX_EXCHANGE_TIMEOUT_NS 5000000000ull

__device__ __forceinline__ u32 pcx_ld_acquire(const u32* ptr) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

// tiles of phase q = [a, b) inside this engine's range [B, E) that signal the
// ticket for this output selection
template <class Ph>
__device__ __forceinline__ int pcx_signalling_tiles(int a, int b, int B, int E) {
    constexpr int F = PCX_FLAGS;
    if (pcx_need_red<Ph>()) {
        const int lo = a > B ? a : B, hi = b < E ? b : E;
        return hi > lo ? hi - lo : 0;
    }
    if ((F & PCX_F_H) || (F & PCX_F_GRAD)) {
        int n = (a >= B && a < E) ? 1 : 0;
        if (b - 1 != a && b - 1 >= B && b - 1 < E) ++n;
        return n;
    }
    return 0;
}

__device__ void pcx_border_map(const PcxParams& p, const int inst, const double* bv,
                               const double* sRS, const bool commit, double* sink_slot)
{
    constexpr int F = PCX_FLAGS;
    constexpr int T = PCX_THREADS;
    double sink = 0.0;
    for (int e = threadIdx.x; e < p.n_border; e += T) {
        const int grp = p.border_grp[e];
        const i64 slot = p.border_slot[e];
        const int k0 = p.border_ptr[e], k1 = p.border_ptr[e + 1];
        double* out;
        if (grp == 0) { if (!(F & PCX_F_C)) continue; out = p.c + (i64)inst * p.num_c; }
        else if (grp == 1) { if (!(F & PCX_F_G)) continue; out = p.gj + (i64)inst * p.nnz_g; }
        else if (grp == 2) { if (!(F & PCX_F_H)) continue; out = p.hs + (i64)inst * p.nnz_h; }
        else if (grp == 3) { if (!(F & PCX_F_J)) continue; out = p.jval + inst; }
        else { if (!(F & PCX_F_GRAD)) continue; out = p.grad + (i64)inst * p.num_x; }
        double acc = 0.0;
        for (int k = k0; k < k1; ++k)
            acc += p.border_coef[k] * bv[p.border_bv[k]] * sRS[p.border_rs[k]];
        if (commit) out[slot] = acc; else sink += acc;
    }
    // the warm-up pass only pulls the map tables into L1/L2: its sums go to a
    // scratch word nobody reads, which keeps the loads alive
    if (!commit) *sink_slot = sink;
}

__device__ __noinline__ void pcx_border(const PcxParams& p, const int inst, double* scratch,
                                        const bool wait_for_tiles = true)
{
    constexpr int F = PCX_FLAGS;
    constexpr int T = PCX_THREADS;
    const int tid = threadIdx.x;
    const double* bvg = p.bv + (i64)inst * p.bv_size;
    const double* x = p.x + (i64)inst * p.num_x;
    __shared__ double sRS[1 + PCX_NUM_PHASES];
    __shared__ double sPt[PCX_NPOINT > 0 ? PCX_NPOINT : 1];
    __shared__ double sMult[1 + PCX_NB];
    // the border-value vector lives in shared memory for the whole pass
    double* bv = scratch + 32;

    if (!p.independent) pcx_grid_dependency_wait();
    // ---- independent of the tiles ------------------------------------------
    for (int a = tid; a < PCX_NPOINT; a += T) {
        const double v = pcx_unscale(p.pt_scal[a], x[p.pt_x[a]], p.pt_scal[PCX_NPOINT + a]);
        sPt[a] = v;
        bv[PCX_BV_PTVAL + a] = v;
    }
    for (int k = tid; k <= PCX_NB; k += T) {
        if (k == 0) {
            const double sg = (p.sigma != nullptr) ? p.sigma[inst] : 1.0;
            sMult[0] = sg * pcx_c_gscal[PCX_GS_W];
        } else {
            sMult[k] = (F & PCX_F_H)
                ? p.lam[(i64)inst * p.num_c + (p.num_c - PCX_NB) + (k - 1)]
                  * pcx_c_gscal[PCX_GS_WB + (k - 1)]
                : 0.0;
        }
    }
    for (int q = tid; q <= PCX_NUM_PHASES; q += T) {
        if (q == 0) { sRS[0] = 1.0; bv[0] = 1.0; continue; }
        const int qq = q - 1;
        const i64* pb = pcx_c_pbase + PCX_PHASE_PBASE(qq);
        const double* ps = pcx_c_pscal + PCX_PHASE_PSCAL(qq);
        const i64 i0 = pb[PCX_PB_T0X], iF = pb[PCX_PB_TFX];
        const double t0 = pcx_unscale(ps[PCX_PHASE_TINFO(qq) + 0], i0 >= 0 ? x[i0] : 0.0,
                                      ps[PCX_PHASE_TINFO(qq) + 1]);
        const double tF = pcx_unscale(ps[PCX_PHASE_TINFO(qq) + 2], iF >= 0 ? x[iF] : 0.0,
                                      ps[PCX_PHASE_TINFO(qq) + 3]);
        sRS[q] = 0.5 * (tF - t0);
    }
    for (int i = 1 + tid; i < PCX_BV_PTVAL; i += T) bv[i] = 0.0;      // reductions (not yet known)
    for (int i = PCX_BV_IRR + tid; i < p.bv_size; i += T) bv[i] = 0.0;
    __syncthreads();
    if (tid == 0)
        pcx_point_eval(sPt, sMult, bv + PCX_BV_PTFN, bv + PCX_BV_PTD1, bv + PCX_BV_PTD2);
    __syncthreads();
    const int mode = p.border_mode;
    const int xlen = (PCX_BV_PTVAL - 1) + (p.bv_size - PCX_BV_IRR);
    double* xb = p.xbuf + (i64)inst * xlen;
    if (mode != 1 && (mode != 3 || p.rank == p.border_rank))
        pcx_border_map(p, inst, bv, sRS, false, scratch + 31);   // warm the map tables

    if (mode == 2) {
        // stage 2 of a sharded evaluation: reductions and end-node values come
        // all-reduced over the ranks
        for (int i = 1 + tid; i < PCX_BV_PTVAL; i += T) bv[i] = xb[i - 1];
        for (int i = PCX_BV_IRR + tid; i < p.bv_size; i += T)
            bv[i] = xb[(PCX_BV_PTVAL - 1) + (i - PCX_BV_IRR)];
        __syncthreads();
        pcx_border_map(p, inst, bv, sRS, true, nullptr);
        return;
    }

    // ---- wait for the tiles that feed the border ------------------------------
    const int tB = p.tile_begin, tE = p.tile_begin + p.tile_count;
    int expected = 0;
#define PCX_CASE(P) expected += pcx_signalling_tiles<PcxPhase<P> >(                      \
        (int)pcx_c_pbase[PCX_PHASE_PBASE(P) + PCX_PB_TILE0],                             \
        (int)pcx_c_pbase[PCX_PHASE_PBASE(P) + PCX_PB_TILE1], tB, tE);
    PCX_FOREACH_PHASE(PCX_CASE)
#undef PCX_CASE
    if (expected > 0 && wait_for_tiles) {
        if (tid == 0) {
            while (pcx_ld_acquire(p.ticket + inst) < (u32)expected) __nanosleep(64);
        }
        __syncthreads();
    }
    PCX_STAMP_B(9);
    // end-node values left by the first/last tiles (L2 reads: this CTA's L1 may
    // hold nothing of them, but stay clear of it anyway)
    for (int i = PCX_BV_IRR + tid; i < p.bv_size; i += T) bv[i] = __ldcg(bvg + i);
    // final reductions: deterministic (fixed stride, fixed shuffle tree)
    for (int q = 0; q < PCX_NUM_PHASES; ++q) {
        bool need = false;
#define PCX_CASE(P) if (q == P) need = pcx_need_red<PcxPhase<P> >();
        PCX_FOREACH_PHASE(PCX_CASE)
#undef PCX_CASE
        const int nred = need ? PCX_PHASE_NRED(q) : 0;
        const i64* pbq = pcx_c_pbase + PCX_PHASE_PBASE(q);
        int t_lo = (int)pbq[PCX_PB_TILE0], t_hi = (int)pbq[PCX_PB_TILE1];
        if (t_lo < tB) t_lo = tB;
        if (t_hi > tE) t_hi = tE;
        for (int k = 0; k < nred; ++k) {
            double acc = 0.0;
            for (int t = t_lo + tid; t < t_hi; t += T)
                acc += __ldcg(p.partials + ((i64)inst * p.num_tiles + t) * p.nred_max + k);
            const double r = pcx_block_sum(acc, scratch);
            if (tid == 0) bv[PCX_PHASE_REDOFF(q) + k] = r;
        }
    }
    __syncthreads();
    if (mode == 1) {
        // stage 1 of a sharded evaluation: publish this rank's share
        for (int i = 1 + tid; i < PCX_BV_PTVAL; i += T) xb[i - 1] = bv[i];
        for (int i = PCX_BV_IRR + tid; i < p.bv_size; i += T)
            xb[(PCX_BV_PTVAL - 1) + (i - PCX_BV_IRR)] = bv[i];
    } else if (mode == 3) {
        // fused exchange over peer memory: write this rank's share into the
        // border rank's buffer, make it visible system-wide, publish the epoch.
        // Shares are double-buffered by epoch parity; before a slot is reused
        // the border rank must have consumed its previous content (epoch - 2).
        const int par = (int)(p.epoch & 1ull);
        if (tid == 0 && p.epoch > 2ull) {
            unsigned long long done, t0 = pcx_globaltimer();
            do {
                done = pcx_ld_acquire_sys(p.peer_done + inst);
                if (done + 2ull < p.epoch) {
                    __nanosleep(200);
                    if (pcx_globaltimer() - t0 > PCX_EXCHANGE_TIMEOUT_NS) { atomicExch(p.status, 1u); break; }
                }
            } while (done + 2ull < p.epoch);
        }
        __syncthreads();
        double* mine = p.peer_xbuf + (((i64)p.rank * 2 + par) * p.batch + inst) * p.bv_size;
        for (int i = 1 + tid; i < PCX_BV_PTVAL; i += T) mine[i - 1] = bv[i];
        for (int i = PCX_BV_IRR + tid; i < p.bv_size; i += T)
            mine[(PCX_BV_PTVAL - 1) + (i - PCX_BV_IRR)] = bv[i];
        __threadfence_system();
        __syncthreads();
        if (tid == 0)
            asm volatile("st.release.sys.global.u64 [%0], %1;"
                         :: "l"(p.peer_flags + (i64)p.rank * p.batch + inst), "l"(p.epoch) : "memory");
        if (p.rank == p.border_rank) {
            // wait for every rank's share of this evaluation (any number of ranks,
            // whatever the CTA size), then sum in rank order (deterministic) and
            // apply the border map
            for (int r = tid; r < p.world; r += T) {
                unsigned long long seen, t0 = pcx_globaltimer();
                do {
                    seen = pcx_ld_acquire_sys(p.peer_flags + (i64)r * p.batch + inst);
                    if (seen < p.epoch) {
                        __nanosleep(200);
                        if (pcx_globaltimer() - t0 > PCX_EXCHANGE_TIMEOUT_NS) { atomicExch(p.status, 1u); break; }
                    }
                } while (seen < p.epoch);
            }
            __syncthreads();
            for (int i = tid; i < xlen; i += T) {
                double acc = 0.0;
                for (int r = 0; r < p.world; ++r)
                    acc += __ldcv(p.peer_xbuf + (((i64)r * 2 + par) * p.batch + inst) * p.bv_size + i);
                const int j = i < PCX_BV_PTVAL - 1 ? 1 + i : PCX_BV_IRR + (i - (PCX_BV_PTVAL - 1));
                bv[j] = acc;
            }
            __syncthreads();
            if (tid == 0)
                asm volatile("st.release.sys.global.u64 [%0], %1;"
                             :: "l"(p.peer_done + inst), "l"(p.epoch) : "memory");
            pcx_border_map(p, inst, bv, sRS, true, nullptr);
        }
    } else {
        pcx_border_map(p, inst, bv, sRS, true, nullptr);
    }
    if (tid == 0) p.ticket[inst] = 0u;
}

// pcx_tile is instantiated for the leaders of the body-sharing groups only
template <int P, bool IS_LEADER = (PCX_PHASE_LEADER(P) == P)> struct PcxTileOf {
    static __device__ __forceinline__ void run(const PcxParams& p, const int tile, const int inst,
                                               const int phase, unsigned char* smem,
                                               PcxTileStatic& ts) {
        pcx_tile<PcxPhase<P> >(p, tile, inst, phase, smem, ts);
    }
};
template <int P> struct PcxTileOf<P, false> {
    static __device__ __forceinline__ void run(const PcxParams&, int, int, int, unsigned char*,
                                               PcxTileStatic&) {}
};

extern "C" __global__ void __launch_bounds__(PCX_THREADS, PCX_MIN_BLOCKS)
PCX_KERNEL_NAME(const __grid_constant__ PcxParams p)
{
    extern __shared__ __align__(16) unsigned char pcx_smem[];
    __shared__ PcxTileStatic pcx_tile_static;
    pcx_grid_launch_dependents();
    // large meshes: the border CTA is dispatched first and works in the shadow
    // of the tiles; small ones: last, so that it never holds a slot a tile of
    // its own instance could use
    int tile = p.border_first ? (int)blockIdx.x - 1 : (int)blockIdx.x;
    const int inst = blockIdx.y;
    // multi-start sweeps of small problems (one tile per instance): the tile's
    // own CTA runs the border pass afterwards -- no second CTA, no ticket wait
    const bool solo = (p.num_tiles == 1 && p.border_mode == 0);
    if (solo) tile = 0;
    if (!solo && tile == (p.border_first ? -1 : p.tile_count)) {
#ifndef PCX_DEBUG_NO_BORDER
        pcx_border(p, inst, reinterpret_cast<double*>(pcx_smem));
        PCX_STAMP_B(6);
#endif
        return;
    }
    // tiles are listed phase by phase: the phase follows from the constant-
    // memory tile ranges, without a dependent global load.  The last tile of a
    // phase trades places with the second one, so that both tiles the border
    // pass may wait for (first and last: end-node values) are dispatched first.
    if (p.reverse) tile = p.tile_count - 1 - tile;
    tile += p.tile_begin;
#if PCX_INTERLEAVE_PHASES
    // Dispatch the phases round-robin instead of one after the other: every phase's
    // copy of the generated body is then executed in every generation of CTAs and stays
    // in L2.  Phase by phase, each phase's first generation (and the first generation
    // of every launch) fetches ~200 KB of code per SM that a launch worth of streamed
    // values has long evicted -- all CTAs missing on the same lines at the same time.
    if (PCX_NUM_PHASES > 1 && p.border_mode == 0) {
        int mmin = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < PCX_NUM_PHASES; ++q) {
            const int n = (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE1]
                        - (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE0];
            mmin = n < mmin ? n : mmin;
        }
        const int i = tile;
        if (i < mmin * PCX_NUM_PHASES) {
            const int q = i % PCX_NUM_PHASES;
            int t0 = 0;
#pragma unroll
            for (int r = 0; r < PCX_NUM_PHASES; ++r)
                if (r == q) t0 = (int)pcx_c_pbase[PCX_PHASE_PBASE(r) + PCX_PB_TILE0];
            tile = t0 + i / PCX_NUM_PHASES;
        } else {
            int j = i - mmin * PCX_NUM_PHASES;
            bool done = false;
#pragma unroll
            for (int q = 0; q < PCX_NUM_PHASES; ++q) {
                const int t0 = (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE0];
                const int left = (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE1] - t0 - mmin;
                if (!done && j < left) { tile = t0 + mmin + j; done = true; }
                if (!done) j -= left;
            }
        }
    }
#endif
    int phase = 0;
#pragma unroll
    for (int q = 1; q < PCX_NUM_PHASES; ++q)
        if (tile >= (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE0]) phase = q;
    if (p.border_mode == 0) {
        const int t_lo = (int)pcx_c_pbase[PCX_PHASE_PBASE(phase) + PCX_PB_TILE0];
        const int t_hi = (int)pcx_c_pbase[PCX_PHASE_PBASE(phase) + PCX_PB_TILE1];
        if (t_hi - t_lo > 2) {
            if (tile == t_lo + 1) tile = t_hi - 1;
            else if (tile == t_hi - 1) tile = t_lo + 1;
        }
    } else {
        // one mesh over several GPUs: this rank runs tiles [B, E).  The first and the
        // last tile of a phase feed the border pass (end-node values) and pay a fence
        // for it; dispatched in index order such a tile can be this rank's LAST one
        // (measured: +25 us behind a 65 us grid).  Every phase-end tile inside the
        // range trades places with one of the range's first tiles (the same
        // transpositions in the same order in every CTA: a permutation).
        const int B = p.tile_begin, E = p.tile_begin + p.tile_count;
        int front = B;
#pragma unroll
        for (int q = 0; q < PCX_NUM_PHASES; ++q) {
            const int F = (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE0];
            const int L = (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE1] - 1;
            if (F >= B && F < E) {                      // first tile of phase q
                if (tile == front) tile = F;
                else if (tile == F) tile = front;
                ++front;
            }
            if (L != F && L >= B && L < E) {            // its last tile
                if (tile == front) tile = L;
                else if (tile == L) tile = front;
                ++front;
            }
        }
        phase = 0;
#pragma unroll
        for (int q = 1; q < PCX_NUM_PHASES; ++q)
            if (tile >= (int)pcx_c_pbase[PCX_PHASE_PBASE(q) + PCX_PB_TILE0]) phase = q;
    }
    // one instantiation per LEADER: the phases that share a body run the leader's tile
    // function with their own (run-time) phase index
    switch (PCX_PHASE_LEADER(phase)) {
#define PCX_CASE(P) case P: PcxTileOf<P>::run(p, tile, inst, phase, pcx_smem, pcx_tile_static); break;
        PCX_FOREACH_PHASE(PCX_CASE)
#undef PCX_CASE
        default: break;
    }
    PCX_STAMP(5);
    if (solo) {
        __threadfence();
        __syncthreads();
        pcx_border(p, inst, reinterpret_cast<double*>(pcx_smem), false);
    }
}
