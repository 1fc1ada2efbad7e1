// pcx_params.h -- kernel parameter block shared by the NVRTC-compiled skeleton
// (pcx_kernels.cuh) and the host library (pcx_api.cu).  Plain C layout.
#pragma once

#define PCX_F_C    1
#define PCX_F_DY   2
#define PCX_F_G    4
#define PCX_F_H    8
#define PCX_F_J    16
#define PCX_F_GRAD 32


typedef long long i64;
typedef unsigned int u32;

// recipe word layout -- keep in sync with pycollo_b200/structure.py
// low word: staged row (9), quadrature-table index (13), constant index (7),
// PREV / PLAIN / SKIP; high word: variable (8), slot inside the variable's
// period (18), node inside the section (5: up to 20 nodes per section)
#define RC_E_BITS 9
#define RC_B_BITS 13
#define RC_C_BITS 7
#define RC_B_SHIFT 9
#define RC_C_SHIFT 22
#define RC_PREV_BIT 29
#define RC_PLAIN_BIT 30
#define RC_SKIP_BIT 31
#define RC_VAR_SHIFT 32
#define RC_LOCAL_SHIFT 40
#define RC_LOCAL_BITS 18
#define RC_M_SHIFT 58
#define RC_M_BITS 5
// halo of the defect-multiplier tile: rows of the previous section (< 20)
#define PCX_LAM_HALO 24

struct PcxParams {
    // per-call
    const double* x;        // (batch, num_x)   scaled iterate x_tilde
    const double* lam;      // (batch, num_c)   constraint multipliers (H)
    const double* sigma;    // (batch)          objective factor (H); null -> 1
    double* c;              // (batch, num_c)
    double* dy;             // (batch, num_dy)
    double* gj;             // (batch, nnz_g)
    double* hs;             // (batch, nnz_h)
    double* jval;           // (batch)
    double* grad;           // (batch, num_x)
    i64 num_x, num_c, num_dy, nnz_g, nnz_h;
    int num_tiles, batch, nvmax, n_border, bv_size, nred_max, btab_len;
    int border_first;       // 1: the border CTA is blockIdx.x == 0 (resident from the start)
    // mesh sharding over GPUs (pcx_set_shard): this engine launches tiles
    // [tile_begin, tile_begin + tile_count) only.  border_mode 0: whole mesh on
    // one GPU; 1: local reductions / end-node values -> xbuf, no border map
    // (stage 1); 2: border CTA only, reads the all-reduced xbuf (stage 2)
    int tile_begin, tile_count, border_mode;
    double* xbuf;           // (batch, xbuf_len): [reductions | end-node values]
    // border_mode 3: the exchange is fused into the kernel over peer memory
    // (NVLink).  Every rank's border CTA writes its share into slot
    // (rank, epoch & 1) of peer_xbuf (world x 2 x batch x bv_size doubles, in the
    // border rank's memory) and then publishes `epoch` in peer_flags[rank]; the
    // border rank waits for all flags, sums the shares in rank order, applies the
    // border map and publishes `epoch` in peer_done[inst].  A writer reuses a
    // slot (same parity, two evaluations later) only after the border rank has
    // consumed it: a rank can run at most two evaluations ahead.
    double* peer_xbuf; unsigned long long* peer_flags; unsigned long long* peer_done;
    unsigned long long epoch;
    // independent != 0: the caller declared that this evaluation neither reads nor
    // overwrites anything the previous evaluation on the stream writes (a sweep over
    // iterates with distinct buffers): the kernel does not wait for that grid to
    // complete before touching its own data (scratch sets rotate on the host side)
    int rank, world, border_rank, independent;
    // reverse != 0: this launch walks the tile list backwards.  Consecutive launches of a
    // multi-wave grid alternate, so that a launch STARTS with the phase whose generated
    // body the previous launch has just been executing (warm instruction caches) instead
    // of the one it left hundreds of microseconds and a gigabyte of streamed values ago.
    int reverse;
    u32* status;            // sticky device-side error word (0 = ok; 1 = exchange timeout)
    // tiles
    const i64* tile_desc;   // 8 per tile: phase,k0,k1,node0,nn,run0,run1,-
    const int* run_slo; const int* run_shi; const int* run_type;
    const i64* run_gbase;   // nvmax per run: slot of the run's first period
    // sections (all phases concatenated; sec_node has K+1 entries per phase)
    const i64* sec_node; const int* sec_order; const double* sec_h;
    const int* sec_type;
    // recipes
    const unsigned long long* recipes; const int* type_var_off;
    const double* btab; const int* order_a_off; const int* order_w_off;
    // scaling-dependent tables (rewritten by pcx_set_scaling)
    const double* pscal; const double* gscal;
    const i64* pbase;
    // reductions / border
    double* partials; u32* ticket; double* bv;
    const int* border_grp; const i64* border_slot; const int* border_ptr;
    const int* border_bv; const int* border_rs; const double* border_coef;
    const i64* err_desc;    // mesh-error pass, 12 per phase: x_off,c_off,N,K,NY,sec_node_off,err_off,sec_off,mmax
    // solution re-fit onto the p+1 mesh (pcx_refit_to_ph): 16 i64 per phase + one
    // global row, 2 doubles per phase (fixed t0 / tF), the per-order Cy / Pu
    // matrices and their offsets (2 ints per order)
    const i64* refit_desc; const double* refit_const;
    const double* refit_tab; const int* refit_off;
    const i64* pt_x;        // x index of each point variable
    const double* pt_scal;  // V then r of each point variable
};

