// Host orchestration of the fixed-grid CDE solve: tiling plan, workspace layout, per-stage launch sequence.
// No host synchronisation anywhere: every call only enqueues work on the caller's stream.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "solve_kernels.cuh"
#include "field_tc.cuh"
#include "hidden_tc.cuh"
#include "persist_tc.cuh"
#include "adaptive_kernels.cuh"

namespace ncde {

static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---- optional per-kernel event timing ------------------------------------------------------------------------
struct ProfRecord { int cls; cudaEvent_t a, b; };
static int g_prof_mask = 0;
static std::vector<ProfRecord> g_prof;

struct ProfScope {
    int cls; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; bool on;
    ProfScope(int c, cudaStream_t s) : cls(c), st(s), on((g_prof_mask >> c) & 1) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(b, st); g_prof.push_back({cls, a, b}); }
    }
};

constexpr size_t kSmemLimit = 232448;  // 227 KB opt-in dynamic shared memory per CTA on sm_100

// ---- TMA descriptors (tensor-core path) -----------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so that the library has no link-time
// dependency on libcuda (it must load on machines without a driver, e.g. for the symbol check of the CPU test suite).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int get_encoder(EncodeTiledFn* fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NCDE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        NCDE_REQUIRE(qres == cudaDriverEntryPointSuccess && ptr, NCDE_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        cached = (EncodeTiledFn)ptr;
    }
    *fn = cached;
    return NCDE_OK;
}
// rank-3 row-major tensor {inner, rows, recs}: element (c, r, k) at base + k*rec_stride + r*row_stride + c*elem
static int make_map(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t rows,
                    uint64_t recs, uint64_t row_stride_bytes, uint64_t rec_stride_bytes, uint32_t box_inner, uint32_t box_rows,
                    CUtensorMapSwizzle sw) {
    EncodeTiledFn enc;
    int rc = get_encoder(&enc);
    if (rc != NCDE_OK) return rc;
    const cuuint64_t dims[3] = {inner, rows, recs ? recs : 1};
    // a single record still needs a legal (multiple of 16) stride
    const cuuint64_t strides[2] = {row_stride_bytes, rec_stride_bytes ? rec_stride_bytes : row_stride_bytes * rows};
    const cuuint32_t box[3] = {box_inner, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    NCDE_REQUIRE(((uintptr_t)base & 15) == 0 && (strides[0] & 15) == 0 && (strides[1] & 15) == 0 && ((uint64_t)box_inner * elem_bytes & 15) == 0,
                 NCDE_ERR_INVALID, "TMA descriptor: misaligned tensor");
    const CUresult r = enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NCDE_REQUIRE(r == CUDA_SUCCESS, NCDE_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return NCDE_OK;
}
// rank-4 bf16 tensor {128 inner, rows, parts, recs} with the 128-byte swizzle, box {64, box_rows, 1, 1} (persistent kernels:
// operand tiles whose (hi, lo) parts and stage records sit at fixed strides)
static int make_map4(CUtensorMap* m, const void* base, uint64_t rows, uint64_t parts, uint64_t recs, uint64_t part_stride_bytes,
                     uint64_t rec_stride_bytes, uint32_t box_rows) {
    EncodeTiledFn enc;
    int rc = get_encoder(&enc);
    if (rc != NCDE_OK) return rc;
    if (recs < 1) recs = 1;
    const cuuint64_t dims[4] = {128, rows, parts, recs};
    if (!part_stride_bytes) part_stride_bytes = 256 * rows;
    if (!rec_stride_bytes) rec_stride_bytes = part_stride_bytes * parts;
    const cuuint64_t strides[3] = {256, part_stride_bytes, rec_stride_bytes};
    const cuuint32_t box[4] = {64, box_rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    NCDE_REQUIRE(((uintptr_t)base & 15) == 0 && (strides[1] & 15) == 0 && (strides[2] & 15) == 0, NCDE_ERR_INVALID,
                 "TMA descriptor: misaligned tensor");
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NCDE_REQUIRE(r == CUDA_SUCCESS, NCDE_ERR_CUDA, "cuTensorMapEncodeTiled (rank 4) failed with code %d", (int)r);
    return NCDE_OK;
}
// The activation / dX records of one solve sit at a fixed stride; launches address them by record index.
struct TcMapSet {
    TcMaps maps;
    const char* base_a; size_t stride_a; int64_t n_a;
    const char* base_x; size_t stride_x; int64_t n_x;
};
constexpr int kNumSMs = 148;

struct Plan {
    int B, Bp, H, C, Cp, F, DF, DFP, Dmax;   // C / Cp: channels of the f(z).dX/dt contraction (1 / 4 when vf != MATMUL)
    int vf, PC;                              // ncde_vf_type; channels of the control path
    int gated;                               // sigmoid-gated final layer: Cp = 2 * (C padded to 2) interleaved columns per h
    int D[NCDE_MAX_LAYERS + 1];
    int Dp4[NCDE_MAX_LAYERS + 1];
    int Hg, S, n_hg, Np, n_bt, Bt, TM;
    int ns;                   // operand parts of the tensor-core tiles: 1 = bf16, 2 = bf16x3 (hi, lo)
    int tc_hid;               // hidden layers on the tensor cores too (fixed-grid bf16 path): bf16 row-major activation records
    size_t abl_off[NCDE_MAX_LAYERS + 1];   // tc_hid: record offset (floats) of the bf16 [Bp][128] input of layer l
    size_t off_Wh, off_bh;    // tc_hid: packed bf16 hidden weights [F][128][128], fp32 biases [F][128]
    int bwd_ew;               // epilogue warps of the tensor-core backward kernel (8 or 16)
    int tc, Npad, KP, CpB;    // tensor-core path: columns per h-group padded to 16, K padded to 128; dX row pitch in smem
    int R, n_rt;
    int n_stages;
    size_t stage_floats;      // saved floats per RK stage
    size_t act_off[NCDE_MAX_LAYERS + 1];
    size_t dx_off, abf_off;
    size_t fwd_smem, bwd_smem, hid_smem;
    size_t off_WT[NCDE_MAX_LAYERS], off_WR[NCDE_MAX_LAYERS], off_bp[NCDE_MAX_LAYERS], off_W3T, off_W3R, off_b3p,
        wpack_floats;
    int ldw[NCDE_MAX_LAYERS], ldi[NCDE_MAX_LAYERS];
    int first_of_slot[NCDE_MAX_LAYERS];  // first layer that uses the same parameters as layer l
    size_t wt_floats, wr_floats;         // contiguous packed hidden weights (k-major / row-major), unique slots only
    int w_in_smem;
    size_t hid_smem_fwd, hid_smem_bwd;
    int wg_tiles, wg_split, wg_rows;     // hidden weight-gradient grid
};

static size_t bwd_smem_floats(int DF, int DFP, int S, int Hg) {
    return (size_t)DF * S + (size_t)S * DFP + (size_t)DF * kChunk + (size_t)kChunk * (DFP + 4) +
           (size_t)S * kGnStride + (size_t)kChunk * S + (size_t)Hg * kChunk;
}
static size_t fwd_smem_floats(int DF, int S) {
    return (size_t)DF * S + (size_t)DF * kChunk + (size_t)(S / 4) * kChunk;
}

// fixed_path: the stored-stage fixed-grid forward / backward (all-tensor-core records allowed);
// fixed_adjoint: the fixed-grid continuous adjoint (same CUDA-core kernels, so evaluate / derivative and gated fields are allowed)
static int make_plan(const ncde_problem_t* p, Plan* pl, bool fixed_path = false, bool fixed_adjoint = false, bool for_bwd = false) {
    const bool fixed_kernels = fixed_path || fixed_adjoint;
    memset(pl, 0, sizeof(*pl));
    const ncde_mlp_t& m = p->mlp;
    NCDE_REQUIRE(p->B >= 1 && p->H >= 1 && p->C >= 1, NCDE_ERR_INVALID, "solve: B, H, C must be positive");
    NCDE_REQUIRE(p->B < (1ll << 30), NCDE_ERR_UNSUPPORTED, "solve: batch too large");
    NCDE_REQUIRE(m.n_layers >= 1 && m.n_layers <= NCDE_MAX_LAYERS, NCDE_ERR_INVALID, "solve: 1..%d layers",
                 NCDE_MAX_LAYERS);
    NCDE_REQUIRE(p->method == NCDE_EULER || p->method == NCDE_RK4_38 || p->method == NCDE_DOPRI5, NCDE_ERR_UNSUPPORTED,
                 "solve: unknown method %d", p->method);
    NCDE_REQUIRE(p->vf_type >= NCDE_VF_MATMUL && p->vf_type <= NCDE_VF_DERIVATIVE, NCDE_ERR_INVALID,
                 "vector_field_type string not recognised");
    pl->vf = p->vf_type; pl->PC = p->C;
    if (pl->vf) {
        // f([z, X(t)]) / f([z, dX/dt(t)]) (torchcde/solver.py:123-126): the control is part of the first layer's input and the
        // final layer is a plain H-wide tanh layer, run by the field kernels as a contraction over ONE channel with dX/dt = 1
        NCDE_REQUIRE(p->precision == NCDE_PREC_FP32, NCDE_ERR_UNSUPPORTED,
                     "solve: vector_field_type evaluate / derivative runs on the fp32 path only");
        NCDE_REQUIRE(p->method != NCDE_DOPRI5, NCDE_ERR_UNSUPPORTED,
                     "solve: vector_field_type evaluate / derivative is implemented for the fixed-grid solvers only");
        NCDE_REQUIRE(p->path.match == nullptr, NCDE_ERR_UNSUPPORTED,
                     "solve: vector_field_type evaluate / derivative is not implemented for gradient-matched paths");
        NCDE_REQUIRE(fixed_kernels, NCDE_ERR_UNSUPPORTED,
                     "solve: vector_field_type evaluate / derivative is implemented for the fixed-grid solvers only");
    }
    pl->B = (int)p->B; pl->H = p->H; pl->C = pl->vf ? 1 : p->C;
    pl->Bp = (int)round_up(p->B, kTcM);
    // channels padded to 4 (float4 epilogues) or 8 (tensor-core path: 16-byte bf16 chunks per h)
    pl->Cp = (int)round_up(pl->C, p->precision != NCDE_PREC_FP32 ? 8 : 4);
    pl->gated = m.W_gate != nullptr;
    if (pl->gated) {
        NCDE_REQUIRE(p->precision == NCDE_PREC_FP32 && fixed_kernels, NCDE_ERR_UNSUPPORTED,
                     "solve: gated vector fields run on the fixed-grid fp32 path only (no adaptive solver)");
        NCDE_REQUIRE(m.n_layers < NCDE_MAX_LAYERS, NCDE_ERR_UNSUPPORTED, "solve: gated vector fields take at most %d layers",
                     NCDE_MAX_LAYERS - 1);
        pl->Cp = 2 * (int)round_up(pl->C, 2);   // (sigmoid, tanh) column pairs
        NCDE_REQUIRE(pl->Cp <= 128, NCDE_ERR_UNSUPPORTED, "solve: gated vector fields support at most 64 channels, got %d", pl->C);
    }
    pl->F = m.n_layers - 1;
    pl->n_stages = p->method == NCDE_RK4_38 ? 4 : (p->method == NCDE_DOPRI5 ? 7 : 1);
    NCDE_REQUIRE(m.in_dim[0] == p->H + (pl->vf ? p->C : 0), NCDE_ERR_INVALID, "solve: first layer must take %d inputs, takes %d",
                 p->H + (pl->vf ? p->C : 0), m.in_dim[0]);
    for (int l = 0; l < m.n_layers; ++l) {
        NCDE_REQUIRE(m.W[l] != nullptr, NCDE_ERR_INVALID, "solve: layer %d has no weight", l);
        NCDE_REQUIRE(m.in_dim[l] >= 1 && m.out_dim[l] >= 1, NCDE_ERR_INVALID, "solve: layer %d has empty shape", l);
        if (l > 0)
            NCDE_REQUIRE(m.in_dim[l] == m.out_dim[l - 1], NCDE_ERR_INVALID, "solve: layer %d input %d != previous output %d",
                         l, m.in_dim[l], m.out_dim[l - 1]);
        NCDE_REQUIRE(m.act[l] >= NCDE_ACT_NONE && m.act[l] <= NCDE_ACT_GATE_IN, NCDE_ERR_INVALID, "solve: bad activation");
        if (m.act[l] == NCDE_ACT_GATE_IN) {
            NCDE_REQUIRE(l < m.n_layers - 1 && m.out_dim[l] == 2 * m.in_dim[l], NCDE_ERR_INVALID,
                         "solve: a gate-in layer must be a hidden layer with out_dim == 2 * in_dim");
            NCDE_REQUIRE(p->precision == NCDE_PREC_FP32 && fixed_kernels, NCDE_ERR_UNSUPPORTED,
                         "solve: gate-in layers (GRU-gated vector fields) run on the fixed-grid fp32 path only");
        }
        pl->D[l] = m.in_dim[l];
        pl->Dp4[l] = (int)round_up(m.in_dim[l], 4);
    }
    NCDE_REQUIRE(m.out_dim[pl->F] == p->H * pl->C, NCDE_ERR_INVALID,
                 "solve: final layer must produce %d outputs, produces %d", p->H * pl->C, m.out_dim[pl->F]);
    NCDE_REQUIRE(m.act[pl->F] == NCDE_ACT_TANH, NCDE_ERR_UNSUPPORTED, "solve: final activation must be tanh");
    pl->DF = pl->D[pl->F];
    pl->DFP = (int)round_up(pl->DF, 16);
    // fp32 kernels take final-layer inputs up to 256 wide (the reference's hyper-parameter search reaches 196,
    // experiments/configurations/configurations.json5:35); the tensor-core tiles are built for K <= 128
    NCDE_REQUIRE(pl->DF <= (p->precision != NCDE_PREC_FP32 ? 128 : 256), NCDE_ERR_UNSUPPORTED,
                 "solve: final-layer input width %d > %d not supported for this precision", pl->DF,
                 p->precision != NCDE_PREC_FP32 ? 128 : 256);
    NCDE_REQUIRE(pl->Cp <= 128, NCDE_ERR_UNSUPPORTED, "solve: %d input channels > 128 not supported", p->C);
    int dmax = pl->Cp;
    for (int l = 0; l <= pl->F; ++l) dmax = pl->Dp4[l] > dmax ? pl->Dp4[l] : dmax;
    NCDE_REQUIRE(dmax <= 1024, NCDE_ERR_UNSUPPORTED, "solve: layer width %d > 1024 not supported", dmax);
    pl->Dmax = dmax;

    pl->tc = p->precision != NCDE_PREC_FP32;
    pl->ns = p->precision == NCDE_PREC_BF16X3 ? 2 : 1;
    if (pl->ns == 2) {
        NCDE_REQUIRE(fixed_path && !pl->vf && !pl->gated, NCDE_ERR_UNSUPPORTED,
                     "solve: precision bf16x3 runs on the fixed-grid solvers (euler / rk4, backprop through the steps) with "
                     "vector_field_type matmul only");
        NCDE_REQUIRE(m.n_layers >= 2, NCDE_ERR_UNSUPPORTED, "solve: precision bf16x3 needs at least one hidden layer");
    }
    if (!pl->tc) {
        // field tiling (fp32 kernels): h-groups x batch tiles
        double best = -1.0;
        const int hg_max = pl->H < 128 / pl->Cp ? pl->H : 128 / pl->Cp;
        for (int hg = hg_max; hg >= 1; --hg) {
            const int S = hg * pl->Cp, NT = S / 4;
            if (bwd_smem_floats(pl->DF, pl->DFP, S, hg) * 4 > kSmemLimit) continue;
            if (NT * (pl->DFP / 16) > kThreads) continue;
            const int n_hg = (int)ceil_div(pl->H, hg);
            int n_bt = kNumSMs / n_hg;
            const int max_bt = (int)ceil_div(pl->B, kChunk);
            n_bt = n_bt < 1 ? 1 : (n_bt > max_bt ? max_bt : n_bt);
            const int TM = NT > 16 ? 8 : 4;
            const int thr = NT * (kChunk / TM);
            const double util = (thr > kThreads ? kThreads : thr) / (double)kThreads;
            const int ctas = n_hg * n_bt;
            const double score = (ctas > kNumSMs ? kNumSMs : ctas) / (double)kNumSMs * util;
            if (score > best + 1e-9) {
                best = score;
                pl->Hg = hg; pl->S = S; pl->n_hg = n_hg; pl->TM = TM;
                pl->Bt = (int)round_up(ceil_div(pl->B, n_bt), kChunk);
                pl->n_bt = (int)ceil_div(pl->B, pl->Bt);
            }
        }
        NCDE_REQUIRE(best > 0, NCDE_ERR_UNSUPPORTED, "solve: no field tiling fits shared memory (C=%d, width=%d)", p->C, pl->DF);
        pl->Npad = pl->S;
        pl->KP = pl->DFP;
        pl->fwd_smem = fwd_smem_floats(pl->DF, pl->S) * 4;
        pl->bwd_smem = bwd_smem_floats(pl->DF, pl->DFP, pl->S, pl->Hg) * 4;
    } else {
        // tensor-core tiling: h-group columns padded to a multiple of 16 (UMMA N), K padded to 64 (one swizzle block)
        // K is padded to 128 so that the weight-gradient MMA (M = K) always runs the M = 128 shape
        pl->KP = kTcKP;
        {
            static const char* ew = getenv("NCDE_BWD_EW");
            pl->bwd_ew = (ew && atoi(ew) == 16) ? 16 : 8;   // measured: 16 epilogue warps are 10 % slower (40.6 vs 36.8 us per launch, cfg 5)
        }
        pl->DFP = pl->KP;  // gradient buffers share the padded K
        pl->CpB = pl->Cp + (((pl->Cp / 4) & 1) ? 0 : 4);   // CpB / 4 odd: per-row LDS.128 of consecutive lanes hit distinct banks
        double best = -1.0;
        const int n_mt = (int)ceil_div(pl->B, kTcM);
        // the padded dX/dt row pitch is dropped (bank conflicts instead of no fit) when shared memory is too tight for it
        for (int attempt = 0; attempt < 2 && best < 0; ++attempt) {
            if (attempt == 1) pl->CpB = pl->Cp;
            static const char* force_hg = getenv("NCDE_TC_HG");   // experiments: fix the number of hidden rows per h-group
            for (int hg = 8; hg >= 1; --hg) {
                if (hg > pl->H) continue;
                if (force_hg && atoi(force_hg) > 0 && hg != atoi(force_hg) && atoi(force_hg) <= pl->H) continue;
                const int npad = (int)round_up(hg * pl->Cp, 16);
                if (npad > 240) continue;  // TMEM: [pre | dW^T] must fit 512 columns in the backward kernel
                if (pl->ns == 2) {
                    // bf16x3 runs on the persistent kernels only: the pass at hand decides (weights are re-packed per pass)
                    if (for_bwd ? ps_bwd_smem_bytes(npad, 2, m.n_layers - 1, 2, hg) > kSmemLimit
                                : ps_fwd_smem_bytes(npad, 2, 1, m.n_layers - 1, 2) > kSmemLimit) continue;
                } else {
                    if (tc_bwd_smem_bytes(npad, pl->CpB, pl->bwd_ew) > kSmemLimit) continue;
                    if (tc_fwd_smem_bytes(npad, pl->CpB) > kSmemLimit) continue;
                }
                const int n_hg = (int)ceil_div(pl->H, hg);
                int n_bt = kNumSMs / n_hg;
                n_bt = n_bt < 1 ? 1 : (n_bt > n_mt ? n_mt : n_bt);
                const int ctas = n_hg * n_bt;
                const double waste = (double)(hg * pl->Cp) / npad;
                const double score = (ctas > kNumSMs ? kNumSMs : ctas) / (double)kNumSMs * waste + 1e-4 * hg;
                if (score > best + 1e-9) {
                    best = score;
                    pl->Hg = hg; pl->S = hg * pl->Cp; pl->Npad = npad; pl->n_hg = n_hg;
                    pl->Bt = (int)round_up(ceil_div(pl->B, n_bt), kTcM);
                    pl->n_bt = (int)ceil_div(pl->B, pl->Bt);
                }
            }
        }
        NCDE_REQUIRE(best > 0, NCDE_ERR_UNSUPPORTED, "solve: no tensor-core tiling fits (C=%d, width=%d)", p->C, pl->DF);
        pl->TM = 8;
        pl->fwd_smem = tc_fwd_smem_bytes(pl->Npad, pl->CpB);
        pl->bwd_smem = tc_bwd_smem_bytes(pl->Npad, pl->CpB, pl->bwd_ew);
    }
    pl->Np = pl->n_hg * pl->Npad;
    if (getenv("NCDE_DEBUG_PLAN"))
        fprintf(stderr, "ncde plan: B=%d H=%d C=%d Cp=%d tc=%d Hg=%d Npad=%d n_hg=%d n_bt=%d Bt=%d fwd_smem=%zu bwd_smem=%zu\n", pl->B, pl->H,
                pl->C, pl->Cp, pl->tc, pl->Hg, pl->Npad, pl->n_hg, pl->n_bt, pl->Bt, pl->fwd_smem, pl->bwd_smem);

    // hidden tiling
    int R = (int)round_up(ceil_div(pl->B, kNumSMs), 4);
    R = R < 4 ? 4 : (R > 32 ? 32 : R);
    while (R > 4 && (size_t)2 * pl->Dmax * R * 4 > 48 * 1024) R -= 4;   // wide layers: fewer rows per CTA instead of failing
    pl->R = R;
    pl->n_rt = (int)ceil_div(pl->B, R);
    pl->hid_smem = (size_t)2 * pl->Dmax * R * 4;
    NCDE_REQUIRE(pl->hid_smem <= 48 * 1024, NCDE_ERR_UNSUPPORTED, "solve: hidden tile too large");
    for (int l = 0; l < pl->F; ++l) {
        pl->first_of_slot[l] = l;
        for (int j = 0; j < l; ++j)
            if (m.slot[j] == m.slot[l] && m.W[j] == m.W[l]) { pl->first_of_slot[l] = j; break; }
        if (pl->first_of_slot[l] != l)
            NCDE_REQUIRE(m.in_dim[l] == m.in_dim[pl->first_of_slot[l]] && m.out_dim[l] == m.out_dim[pl->first_of_slot[l]],
                         NCDE_ERR_INVALID, "solve: layers sharing a slot must have the same shape");
    }

    // saved-per-stage layout
    size_t off = 0;
    for (int l = 0; l <= pl->F; ++l) { pl->act_off[l] = off; off += (size_t)pl->Dp4[l] * pl->Bp; }
    pl->dx_off = off; off += (size_t)pl->Cp * pl->Bp;
    pl->abf_off = off;
    if (pl->tc) off += (size_t)pl->Bp * pl->KP / 2;  // bf16 copy of the final-layer input
    pl->stage_floats = off;
    {
        int wmax = pl->H;
        for (int l = 0; l <= pl->F; ++l) wmax = pl->D[l] > wmax ? pl->D[l] : wmax;
        static const bool enabled = getenv("NCDE_NO_TC_HIDDEN") == nullptr;   // NCDE_NO_TC_HIDDEN=1: CUDA-core hidden layers (A/B runs)
        pl->tc_hid = fixed_path && pl->tc && wmax <= 128 && pl->F <= 3 && enabled;
    }
    if (pl->tc_hid) {
        // records of the all-tensor-core path: bf16 [Bp][128] input of every layer (the last one feeds the final layer), dX/dt
        off = 0;
        for (int l = 0; l <= pl->F; ++l) { pl->abl_off[l] = off; off += (size_t)pl->ns * pl->Bp * 64; }
        pl->abf_off = pl->abl_off[pl->F];
        pl->dx_off = off; off += (size_t)pl->Cp * pl->Bp;
        pl->stage_floats = off;
    }

    // packed weights
    off = 0;
    for (int l = 0; l < pl->F; ++l) {  // k-major copies, unique slots, contiguous
        pl->ldw[l] = (int)round_up(m.out_dim[l], 4);
        pl->ldi[l] = (int)round_up(m.in_dim[l], 4);
        if (pl->first_of_slot[l] == l) { pl->off_WT[l] = off; off += (size_t)pl->D[l] * pl->ldw[l]; }
        else pl->off_WT[l] = pl->off_WT[pl->first_of_slot[l]];
    }
    pl->wt_floats = off;
    off = round_up(off, 64);
    const size_t wr_begin = off;
    for (int l = 0; l < pl->F; ++l) {  // row-major copies for the backward pass
        if (pl->first_of_slot[l] == l) { pl->off_WR[l] = off; off += (size_t)m.out_dim[l] * pl->ldi[l]; }
        else pl->off_WR[l] = pl->off_WR[pl->first_of_slot[l]];
    }
    pl->wr_floats = off - wr_begin;
    off = round_up(off, 64);
    for (int l = 0; l < pl->F; ++l) { pl->off_bp[l] = off; off += round_up(pl->ldw[l], 64); }
    {
        const size_t wmax = pl->wt_floats > pl->wr_floats ? pl->wt_floats : pl->wr_floats;
        pl->w_in_smem = pl->F > 0 && (pl->hid_smem + wmax * 4) <= 200 * 1024;
        pl->hid_smem_fwd = pl->hid_smem + (pl->w_in_smem ? round_up(pl->wt_floats, 4) * 4 : 0);
        const size_t acts_bytes = (size_t)pl->F * pl->Dmax * R * 4;
        if (pl->hid_smem + acts_bytes + wmax * 4 > 200 * 1024) pl->w_in_smem = 0;
        pl->hid_smem_fwd = pl->hid_smem + (pl->w_in_smem ? round_up(pl->wt_floats, 4) * 4 : 0);
        pl->hid_smem_bwd = pl->hid_smem + acts_bytes + (pl->w_in_smem ? round_up(pl->wr_floats, 4) * 4 : 0);
    }
    {
        int tiles = 0;
        for (int l = 0; l < pl->F; ++l)
            if (pl->first_of_slot[l] == l) tiles += (int)(ceil_div(m.out_dim[l], kWgTile) * ceil_div(m.in_dim[l], kWgTile));
        pl->wg_tiles = tiles;
        int split = tiles > 0 ? kNumSMs / tiles : 1;
        const int max_split = (int)ceil_div(pl->B, kWgRows);
        split = split < 1 ? 1 : (split > max_split ? max_split : split);
        pl->wg_rows = (int)round_up(ceil_div(pl->B, split), kWgRows);
        pl->wg_split = (int)ceil_div(pl->B, pl->wg_rows);
    }
    // fp32 path: W3T [DF][Np] + W3R [Np][DFP]; tensor-core path: bf16 [Np][KP] stored in the W3T slot
    pl->off_W3T = off; off += round_up((size_t)pl->DF * pl->Np, 64);
    pl->off_W3R = off; off += round_up((size_t)pl->Np * pl->DFP, 64);
    pl->off_b3p = off; off += round_up(pl->Np, 64);
    if (pl->ns == 2) { off = round_up(off, 64); pl->off_W3T = off; off += round_up((size_t)2 * pl->Np * 64, 64); }   // bf16x3: [2][Np][128] bf16
    pl->off_Wh = off; off += (size_t)pl->ns * pl->F * 128 * 64;
    pl->off_bh = off; off += (size_t)pl->F * 128;
    pl->wpack_floats = off;
    return NCDE_OK;
}

// Tiling of the path-gradient launch: field_fwd_kernel on the problem with h and c exchanged (see
// pack_final_swapped_kernel): "hidden rows" = the C channels, "channels" = the H hidden rows padded to 4.
struct SwapPlan {
    int Hp, Hg, S, n_hg, Np, n_bt, Bt, TM;
    size_t smem, w3t_floats, b3p_floats;
};
static int make_swap_plan(const Plan& pl, SwapPlan* sp) {
    memset(sp, 0, sizeof(*sp));
    sp->Hp = (int)round_up(pl.H, 4);
    NCDE_REQUIRE(sp->Hp <= 128, NCDE_ERR_UNSUPPORTED, "solve_bwd: path gradients need hidden width <= 128, got %d", pl.H);
    double best = -1.0;
    const int hg_max = pl.C < 128 / sp->Hp ? pl.C : 128 / sp->Hp;
    for (int hg = hg_max; hg >= 1; --hg) {
        const int S = hg * sp->Hp, NT = S / 4;
        if (fwd_smem_floats(pl.DF, S) * 4 > kSmemLimit) continue;
        const int n_hg = (int)ceil_div(pl.C, hg);
        int n_bt = kNumSMs / n_hg;
        const int max_bt = (int)ceil_div(pl.B, kChunk);
        n_bt = n_bt < 1 ? 1 : (n_bt > max_bt ? max_bt : n_bt);
        const int TM = NT > 16 ? 8 : 4;
        const int thr = NT * (kChunk / TM);
        const double util = (thr > kThreads ? kThreads : thr) / (double)kThreads;
        const int ctas = n_hg * n_bt;
        const double score = (ctas > kNumSMs ? kNumSMs : ctas) / (double)kNumSMs * util;
        if (score > best + 1e-9) {
            best = score;
            sp->Hg = hg; sp->S = S; sp->n_hg = n_hg; sp->TM = TM;
            sp->Bt = (int)round_up(ceil_div(pl.B, n_bt), kChunk);
            sp->n_bt = (int)ceil_div(pl.B, sp->Bt);
        }
    }
    NCDE_REQUIRE(best > 0, NCDE_ERR_UNSUPPORTED, "solve_bwd: no path-gradient tiling fits shared memory (H=%d, width=%d)", pl.H, pl.DF);
    sp->Np = sp->n_hg * sp->S;
    sp->smem = fwd_smem_floats(pl.DF, sp->S) * 4;
    sp->w3t_floats = round_up((size_t)pl.DF * sp->Np, 64);
    sp->b3p_floats = round_up(sp->Np, 64);
    return NCDE_OK;
}

struct Carver {
    char* base; size_t used, cap;
    float* take(size_t floats) {
        size_t bytes = round_up(floats * 4, 256);
        float* p = (float*)(base + used);
        used += bytes;
        return p;
    }
};

static size_t fwd_workspace_floats(const Plan& pl, int need_saved_scratch) {
    size_t per = 256 / 4;  // alignment slack per buffer
    size_t n = pl.wpack_floats + per;
    n += (size_t)(2 + pl.n_stages) * ((size_t)pl.H * pl.Bp + per);
    if (need_saved_scratch) n += pl.stage_floats + per;
    if (pl.vf) n += 4 * (size_t)pl.Bp + per;   // the constant "dX/dt" = (1, 0, 0, 0) of the one-channel contraction
    return n;
}
static size_t ps_workspace_floats(const Plan& pl, int64_t n_steps, bool bwd);
static size_t fwd_workspace_extra_floats(const Plan& pl, int64_t n_steps, int need_saved_scratch) {
    size_t per = 256 / 4;
    size_t n = (size_t)n_steps * pl.n_stages + per;
    if (pl.tc_hid) n += ps_workspace_floats(pl, n_steps, false);                                  // device copy of the stage times
    if (need_saved_scratch) n += (size_t)n_steps * pl.n_stages * (pl.vf ? pl.PC : pl.Cp) * pl.Bp + per;  // dX/dt (vf: X or dX/dt) of every stage
    return n;
}
static size_t bwd_workspace_floats(const Plan& pl, int64_t n_steps, const SwapPlan* sp = nullptr) {
    size_t per = 256 / 4;
    size_t n = pl.wpack_floats + per;
    n += (size_t)(1 + pl.n_stages) * (round_up(pl.H, 4) * (size_t)pl.Bp + per);
    if (pl.vf) n += 4 * (size_t)pl.Bp + per;
    if (sp)   // path gradient: swapped final-layer pack, dL/d(dX/dt) of every stage of one step
        n += sp->w3t_floats + sp->b3p_floats + 2 * per + (size_t)pl.n_stages * ((size_t)pl.C * pl.Bp + per);
    n += (size_t)pl.n_hg * pl.Bp * pl.DFP + per;
    for (int l = 0; l < pl.F; ++l) n += (size_t)pl.n_stages * ((size_t)pl.Dp4[l + 1] * pl.Bp + per);
    n += (size_t)pl.n_bt * pl.Np * pl.DFP + per;
    n += (size_t)pl.n_bt * pl.Np + per;
    n += (size_t)pl.wg_split * (pl.wr_floats + 64 * pl.F) + (size_t)pl.wg_split * pl.F * 1024 + 2 * per;
    if (pl.tc_hid)   // dpre records of every stage and hidden layer (consumed by tc_hidden_wgrad), weight-gradient accumulators
        n += (size_t)(n_steps > 0 ? n_steps : 1) * pl.n_stages * pl.F * pl.ns * pl.Bp * 64 + per + (size_t)pl.F * (128 * 128 + 128) + 2 * per +
             (size_t)(pl.Bp / 128) + per + ps_workspace_floats(pl, n_steps, true);
    return n;
}

static int pack_weights(const ncde_problem_t* p, const Plan& pl, float* wpack, int with_rowmajor, cudaStream_t st,
                        int64_t* launches) {
    const ncde_mlp_t& m = p->mlp;
    for (int l = 0; l < pl.F; ++l) {
        int n = pl.D[l] * pl.ldw[l];
        const int nr = m.out_dim[l] * pl.ldi[l];
        n = n > nr ? n : nr;
        pack_hidden_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
            m.W[l], m.bias[l], wpack + pl.off_WT[l], wpack + pl.off_bp[l], with_rowmajor ? wpack + pl.off_WR[l] : nullptr,
            m.out_dim[l], pl.D[l], pl.ldw[l], pl.ldi[l]);
        ++*launches;
    }
    if (pl.tc_hid) {
        for (int l = 0; l < pl.F; ++l) {
            pack_hidden_bf16_kernel<<<64, 256, 0, st>>>(m.W[l], m.bias[l], (__nv_bfloat16*)(wpack + pl.off_Wh) + (size_t)l * 128 * 128,
                                                        wpack + pl.off_bh + (size_t)l * 128, m.out_dim[l], m.in_dim[l]);
            ++*launches;
        }
    }
    if (pl.tc) {
        const int64_t n = (int64_t)pl.Np * pl.KP;
        pack_final_bf16_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
            m.W[pl.F], m.bias[pl.F], (__nv_bfloat16*)(wpack + pl.off_W3T), wpack + pl.off_b3p, pl.H, pl.C, pl.Cp, pl.Hg,
            pl.n_hg, pl.Npad, pl.KP, pl.DF);
    } else {
        const int64_t n = (int64_t)pl.Np * pl.DFP;
        pack_final_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
            m.W[pl.F], m.bias[pl.F], m.W_gate, m.bias_gate, wpack + pl.off_W3T, with_rowmajor ? wpack + pl.off_W3R : nullptr,
            wpack + pl.off_b3p, pl.H, pl.C, pl.Cp, pl.DF, pl.DFP, pl.Np);
    }
    ++*launches;
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

static int validate_grid(const ncde_problem_t* p, const Plan& pl) {
    const ncde_fixed_grid_t& g = p->grid;
    NCDE_REQUIRE(g.n_steps >= 0 && g.n_out >= 1, NCDE_ERR_INVALID, "solve: empty grid");
    NCDE_REQUIRE(g.n_steps == 0 || (g.stage_t && g.dt), NCDE_ERR_INVALID, "solve: grid arrays missing");
    NCDE_REQUIRE(g.n_out == 1 || (g.out_step && g.out_mode && g.out_slope), NCDE_ERR_INVALID, "solve: output map missing");
    int64_t prev = 0;
    for (int64_t j = 1; j < g.n_out; ++j) {
        NCDE_REQUIRE(g.out_step[j] >= prev && g.out_step[j] < g.n_steps, NCDE_ERR_INVALID,
                     "solve: out_step must be non-decreasing and inside the grid");
        NCDE_REQUIRE(g.out_mode[j] >= 0 && g.out_mode[j] <= 2, NCDE_ERR_INVALID, "solve: bad out_mode");
        prev = g.out_step[j];
    }
    NCDE_REQUIRE(p->path.K >= 2 && p->path.knots && p->path.coeffs, NCDE_ERR_INVALID, "solve: bad path");
    NCDE_REQUIRE(p->path.kind == NCDE_PATH_LINEAR || p->path.kind == NCDE_PATH_CUBIC, NCDE_ERR_INVALID, "solve: bad path kind");
    (void)pl;
    return NCDE_OK;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) NCDE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // Ask for exactly the shared-memory carve-out this kernel needs.  Without the hint the driver may leave the split of
    // the previous kernel in place (the tensor-core kernels take 228 KB), which shrinks L1 for the latency-bound hidden-layer
    // kernels that follow them.
    static const bool hint = getenv("NCDE_NO_CARVEOUT_HINT") == nullptr;
    if (hint) {
        int pct = (int)((bytes + 1024 + 2327) * 100 / 233472);
        pct = pct > 100 ? 100 : pct;
        NCDE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    return NCDE_OK;
}

static void fill_tc_args(TcFieldArgs& ta, const Plan& pl, const float* wpack) {
    memset(&ta, 0, sizeof(ta));
    ta.B = pl.B; ta.Bp = pl.Bp; ta.H = pl.H; ta.Cp = pl.Cp; ta.Hg = pl.Hg; ta.n_hg = pl.n_hg; ta.Npad = pl.Npad;
    ta.KP = pl.KP; ta.DF = pl.DF; ta.Bt = pl.Bt; ta.DFP = pl.DFP; ta.CpB = pl.CpB;
    ta.Wbf = (const __nv_bfloat16*)(wpack + pl.off_W3T);
    ta.b3 = wpack + pl.off_b3p;
}

// Descriptors of one solve: weights, and the activation (bf16 [B][KP]) / dX (fp32 [B][Cp]) records at their strides.
static int build_tc_maps(const Plan& pl, const float* wpack, TcMapSet* ms, const void* base_a, size_t stride_a_floats,
                         int64_t n_a, const void* base_x, size_t stride_x_floats, int64_t n_x) {
    memset(ms, 0, sizeof(*ms));
    ms->base_a = (const char*)base_a; ms->stride_a = stride_a_floats * 4; ms->n_a = n_a < 1 ? 1 : n_a;
    ms->base_x = (const char*)base_x; ms->stride_x = stride_x_floats * 4; ms->n_x = n_x < 1 ? 1 : n_x;
    int rc = make_map(&ms->maps.W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpack + pl.off_W3T, kTcKP, (uint64_t)pl.n_hg * pl.Npad, 1,
                      kTcKP * 2, 0, 64, (uint32_t)pl.Npad, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != NCDE_OK) return rc;
    // rows beyond the batch are zero-filled by TMA (padding must not reach the weight-gradient reduction)
    rc = make_map(&ms->maps.A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base_a, kTcKP, (uint64_t)pl.B, (uint64_t)ms->n_a, kTcKP * 2,
                  ms->n_a > 1 ? ms->stride_a : 0, 64, kTcM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != NCDE_OK) return rc;
    return make_map(&ms->maps.X, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base_x, (uint64_t)pl.Cp, (uint64_t)pl.B, (uint64_t)ms->n_x,
                    (uint64_t)pl.Cp * 4, ms->n_x > 1 ? ms->stride_x : 0, (uint32_t)pl.CpB, kTcM, CU_TENSOR_MAP_SWIZZLE_NONE);
}
// record indices of this launch from the pointers the call site selected (ta.abf, ta.dXT)
static int tc_records(TcFieldArgs& ta, const TcMapSet& ms) {
    const size_t oa = (const char*)ta.abf - ms.base_a, ox = (const char*)ta.dXT - ms.base_x;
    const int64_t ra = ms.n_a > 1 ? (int64_t)(oa / ms.stride_a) : 0, rx = ms.n_x > 1 ? (int64_t)(ox / ms.stride_x) : 0;
    NCDE_REQUIRE((const char*)ta.abf >= ms.base_a && ra < ms.n_a && oa == (size_t)ra * (ms.n_a > 1 ? ms.stride_a : 0), NCDE_ERR_INVALID,
                 "tensor-core launch: activation record outside the descriptor");
    NCDE_REQUIRE((const char*)ta.dXT >= ms.base_x && rx < ms.n_x && ox == (size_t)rx * (ms.n_x > 1 ? ms.stride_x : 0), NCDE_ERR_INVALID,
                 "tensor-core launch: dX record outside the descriptor");
    ta.rec_a = (int)ra; ta.rec_x = (int)rx;
    return NCDE_OK;
}
static int launch_tc_fwd(const Plan& pl, TcFieldArgs& ta, const TcMapSet& ms, cudaStream_t st) {
    int rc = tc_records(ta, ms);
    if (rc != NCDE_OK) return rc;
    NCDE_CUDA_OK(launch_pdl(tc_field_fwd_kernel, dim3(pl.n_hg, pl.n_bt), dim3(kTcThreads), pl.fwd_smem, st, ta, ms.maps));
    return NCDE_OK;
}
// descriptors of the tensor-core hidden layers: packed weights and the bf16 activation records (rec0 = first record)
static int build_hidden_maps(const Plan& pl, const float* wpack, TcHiddenMaps* hm, const float* rec0, size_t rec_stride_floats,
                             int64_t n_rec) {
    memset(hm, 0, sizeof(*hm));
    int rc = NCDE_OK;
    if (pl.F > 0)
        rc = make_map(&hm->W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpack + pl.off_Wh, 128, 128, (uint64_t)pl.F, 256, 128 * 256, 64, 128,
                      CU_TENSOR_MAP_SWIZZLE_128B);
    for (int l = 0; l <= pl.F && rc == NCDE_OK; ++l)
        rc = make_map(&hm->act[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rec0 + pl.abl_off[l], 128, (uint64_t)pl.B,
                      (uint64_t)(n_rec < 1 ? 1 : n_rec), 256, n_rec > 1 ? rec_stride_floats * 4 : 0, 64, kTcM, CU_TENSOR_MAP_SWIZZLE_128B);
    return rc;
}

static int launch_tc_bwd(const Plan& pl, TcFieldArgs& ta, const TcMapSet& ms, cudaStream_t st) {
    int rc = tc_records(ta, ms);
    if (rc != NCDE_OK) return rc;
    if (pl.bwd_ew == 16)
        NCDE_CUDA_OK(launch_pdl(tc_field_bwd_kernel<16>, dim3(pl.n_hg, pl.n_bt), dim3(16 * 32 + 32), pl.bwd_smem, st, ta, ms.maps));
    else
        if (ta.gk_early) {   // experimental variant (NCDE_GK_EARLY=1), separate instantiation so that the default kernel's code is untouched
            static bool opted = false;
            if (!opted) { const int rc_o = opt_in_smem(tc_field_bwd_kernel<8, true>, pl.bwd_smem); if (rc_o != NCDE_OK) return rc_o; opted = true; }
            NCDE_CUDA_OK(launch_pdl(tc_field_bwd_kernel<8, true>, dim3(pl.n_hg, pl.n_bt), dim3(8 * 32 + 32), pl.bwd_smem, st, ta, ms.maps));
        } else
        NCDE_CUDA_OK(launch_pdl(tc_field_bwd_kernel<8>, dim3(pl.n_hg, pl.n_bt), dim3(8 * 32 + 32), pl.bwd_smem, st, ta, ms.maps));
    return NCDE_OK;
}

static void fill_field_args(FieldArgs& fa, const Plan& pl, const float* wpack) {
    memset(&fa, 0, sizeof(fa));
    fa.B = pl.B; fa.Bp = pl.Bp; fa.H = pl.H; fa.Cp = pl.Cp; fa.DF = pl.DF; fa.DFP = pl.DFP; fa.S = pl.S; fa.Hg = pl.Hg;
    fa.n_hg = pl.n_hg; fa.Np = pl.Np; fa.Bt = pl.Bt;
    fa.W3T = wpack + pl.off_W3T; fa.W3R = wpack + pl.off_W3R; fa.b3p = wpack + pl.off_b3p;
}

// fp32 final-layer launches of the fixed-grid path (plain or sigmoid-gated epilogue)
static int opt_in_field(const Plan& pl, bool backward) {
    if (backward) {
        if (pl.gated) return pl.TM == 8 ? opt_in_smem(field_bwd_kernel<8, true>, pl.bwd_smem) : opt_in_smem(field_bwd_kernel<4, true>, pl.bwd_smem);
        return pl.TM == 8 ? opt_in_smem(field_bwd_kernel<8>, pl.bwd_smem) : opt_in_smem(field_bwd_kernel<4>, pl.bwd_smem);
    }
    if (pl.gated) return pl.TM == 8 ? opt_in_smem(field_fwd_kernel<8, true>, pl.fwd_smem) : opt_in_smem(field_fwd_kernel<4, true>, pl.fwd_smem);
    return pl.TM == 8 ? opt_in_smem(field_fwd_kernel<8>, pl.fwd_smem) : opt_in_smem(field_fwd_kernel<4>, pl.fwd_smem);
}
static cudaError_t launch_field(const Plan& pl, const FieldArgs& fa, bool backward, cudaStream_t st) {
    const dim3 fg(pl.n_hg, pl.n_bt);
    if (backward) {
        if (pl.gated)
            return pl.TM == 8 ? launch_pdl(field_bwd_kernel<8, true>, fg, dim3(kThreads), pl.bwd_smem, st, fa)
                              : launch_pdl(field_bwd_kernel<4, true>, fg, dim3(kThreads), pl.bwd_smem, st, fa);
        return pl.TM == 8 ? launch_pdl(field_bwd_kernel<8>, fg, dim3(kThreads), pl.bwd_smem, st, fa)
                          : launch_pdl(field_bwd_kernel<4>, fg, dim3(kThreads), pl.bwd_smem, st, fa);
    }
    if (pl.gated)
        return pl.TM == 8 ? launch_pdl(field_fwd_kernel<8, true>, fg, dim3(kThreads), pl.fwd_smem, st, fa)
                          : launch_pdl(field_fwd_kernel<4, true>, fg, dim3(kThreads), pl.fwd_smem, st, fa);
    return pl.TM == 8 ? launch_pdl(field_fwd_kernel<8>, fg, dim3(kThreads), pl.fwd_smem, st, fa)
                      : launch_pdl(field_fwd_kernel<4>, fg, dim3(kThreads), pl.fwd_smem, st, fa);
}


// ---------------------------------------------------------------------------------------------------------------
// persistent whole-pass kernels (persist_tc.cuh): grid shape, eligibility, packing, descriptors
// ---------------------------------------------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // every CTA must be resident: the roles wait for each other
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct PsPlan { int n_mt, n_part, n_field, n_hid, NA, NX; size_t smem; };

static int device_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = kNumSMs;
    }
    return sms;
}

// every output is the state at the end of a step, at most one per step (online outputs on the grid, terminal output)
static bool ps_outputs_on_grid(const ncde_fixed_grid_t& g) {
    int64_t prev = -1;
    for (int64_t j = 1; j < g.n_out; ++j) {
        if (g.out_mode[j] != 1 || g.out_step[j] == prev) return false;
        prev = g.out_step[j];
    }
    return true;
}

static bool ps_eligible(const ncde_problem_t* p, const Plan& pl, bool bwd, PsPlan* pp) {
    static const bool disabled = getenv("NCDE_NO_PERSIST") != nullptr;
    if ((disabled && pl.ns == 1) || !pl.tc_hid || pl.F < 1 || p->grid.n_steps < 1 || pl.vf || pl.gated) return false;
    if (!ps_outputs_on_grid(p->grid)) return false;
    const int sms = device_sm_count();
    pp->n_mt = pl.Bp / kTcM;
    pp->n_hid = pp->n_mt < 16 ? pp->n_mt : 16;
    while (pp->n_hid > 1 && pl.n_hg + pp->n_hid > sms) --pp->n_hid;
    if (pl.n_hg + pp->n_hid > sms) return false;
    pp->n_part = (sms - pp->n_hid) / pl.n_hg;
    if (pp->n_part > pp->n_mt) pp->n_part = pp->n_mt;
    pp->n_field = pp->n_part * pl.n_hg;
    // shared memory: operand tiles first, then as many dX/dt chunk slots as fit (at least 2; more = the loader runs further ahead)
    pp->NA = 1;
    int nch = (pl.Cp + 31) / 32;
    if (bwd && ps_bwd_narrow(pl.Hg)) nch = (pl.Cp + 15) / 16;
    const int want = 2 * nch < kPsMaxSlots ? 2 * nch : kPsMaxSlots;     // two whole tiles' worth is as far ahead as it is useful to run
    if (bwd) {
        pp->NX = 2;
        while (pp->NX < want && ps_bwd_smem_bytes(pl.Npad, pl.ns, pl.F, pp->NX + 1, pl.Hg) <= kSmemLimit) ++pp->NX;
        pp->smem = ps_bwd_smem_bytes(pl.Npad, pl.ns, pl.F, pp->NX, pl.Hg);
    } else {
        pp->NX = 2;
        const int need = nch < kPsMaxSlots ? nch : kPsMaxSlots;        // one tile's worth of slots before a second activation buffer
        while (pp->NX < need && ps_fwd_smem_bytes(pl.Npad, pl.ns, 1, pl.F, pp->NX + 1) <= kSmemLimit) ++pp->NX;
        if (ps_fwd_smem_bytes(pl.Npad, pl.ns, 2, pl.F, pp->NX) <= kSmemLimit) pp->NA = 2;
        while (pp->NX < want && ps_fwd_smem_bytes(pl.Npad, pl.ns, pp->NA, pl.F, pp->NX + 1) <= kSmemLimit) ++pp->NX;
        pp->smem = ps_fwd_smem_bytes(pl.Npad, pl.ns, pp->NA, pl.F, pp->NX);
    }
    return pp->smem <= kSmemLimit && (int64_t)p->grid.n_steps * pl.n_stages * pp->n_mt < (1ll << 30);
}
// synchronisation words + device copies of dt / emit slots, in floats
static size_t ps_workspace_floats(const Plan& pl, int64_t n_steps, bool bwd) {
    size_t n = 2 * (size_t)(n_steps > 0 ? n_steps : 1) + 2 * (size_t)(pl.Bp / kTcM) + 4 * 64;
    if (bwd) n += (size_t)(pl.Bp / kTcM) * 128 * 128 + 64 + (size_t)pl.H * pl.Bp + 64;   // dA^T tiles, second gy buffer
    return n;
}

static int ps_pack(const ncde_problem_t* p, const Plan& pl, float* wpack, cudaStream_t st, int64_t* launches) {
    const ncde_mlp_t& m = p->mlp;
    for (int l = 0; l < pl.F; ++l) {
        ps_pack_hidden_kernel<<<64, 256, 0, st>>>(m.W[l], m.bias[l], (__nv_bfloat16*)(wpack + pl.off_Wh) + (size_t)l * pl.ns * 128 * 128,
                                                  wpack + pl.off_bh + (size_t)l * 128, m.out_dim[l], m.in_dim[l], pl.ns);
        ++*launches;
    }
    const int64_t n = (int64_t)pl.Np * 128;
    ps_pack_final_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(m.W[pl.F], m.bias[pl.F], (__nv_bfloat16*)(wpack + pl.off_W3T),
                                                                     wpack + pl.off_b3p, pl.H, pl.C, pl.Cp, pl.Hg, pl.n_hg, pl.Npad, pl.DF, pl.ns);
    ++*launches;
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

// descriptors: packed weights, the bf16 activation records (rec0 = first record, stride in floats), dpre records (backward)
static int ps_build_maps(const Plan& pl, const float* wpack, PsMaps* pm, const float* rec0, size_t rec_stride_floats, int64_t n_rec,
                         const float* dpre0, const float* dx0, size_t dx_stride_floats, int64_t n_dx, int x_pitch = kPsXPitch) {
    memset(pm, 0, sizeof(*pm));
    int rc = make_map(&pm->W3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpack + pl.off_W3T, 128, (uint64_t)pl.Np, (uint64_t)pl.ns, 256,
                      (uint64_t)pl.Np * 256, 64, (uint32_t)pl.Npad, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc == NCDE_OK) rc = make_map4(&pm->Wh, wpack + pl.off_Wh, 128, (uint64_t)pl.ns, (uint64_t)pl.F, 128 * 256, (uint64_t)pl.ns * 128 * 256, 128);
    for (int l = 0; l <= pl.F && rc == NCDE_OK; ++l)
        rc = make_map4(&pm->act[l], rec0 + pl.abl_off[l], (uint64_t)pl.B, (uint64_t)pl.ns, (uint64_t)n_rec, (uint64_t)pl.Bp * 256,
                       n_rec > 1 ? rec_stride_floats * 4 : 0, kTcM);
    if (rc == NCDE_OK)   // dX/dt records, row-major [B][Cp] fp32; one box = 128 rows x (32 + 4) channels, rows / channels beyond the tensor read 0
        rc = make_map(&pm->X, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx0, (uint64_t)pl.Cp, (uint64_t)pl.B, (uint64_t)(n_dx < 1 ? 1 : n_dx),
                      (uint64_t)pl.Cp * 4, n_dx > 1 ? dx_stride_floats * 4 : 0, (uint32_t)x_pitch, kTcM, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc == NCDE_OK && dpre0)
        rc = make_map4(&pm->dpre, dpre0, (uint64_t)pl.B, (uint64_t)pl.ns, (uint64_t)n_rec * pl.F, (uint64_t)pl.Bp * 256,
                       (uint64_t)pl.ns * pl.Bp * 256, kTcM);
    return rc;
}

static void ps_fill_args(PsArgs& a, const ncde_problem_t* p, const Plan& pl, const PsPlan& pp, const float* wpack) {
    memset(&a, 0, sizeof(a));
    a.B = pl.B; a.Bp = pl.Bp; a.H = pl.H; a.Cp = pl.Cp; a.Hg = pl.Hg; a.n_hg = pl.n_hg; a.Npad = pl.Npad; a.F = pl.F;
    a.n_mt = pp.n_mt; a.n_part = pp.n_part; a.n_field = pp.n_field; a.n_hid = pp.n_hid;
    a.NS = pl.n_stages; a.n_steps = (int)p->grid.n_steps; a.method = p->method; a.NA = pp.NA; a.NX = pp.NX;
    for (int l = 0; l < pl.F; ++l) a.act[l] = p->mlp.act[l];
    a.b3 = wpack + pl.off_b3p;
    a.bias_h = wpack + pl.off_bh;
    for (int l = 0; l <= pl.F; ++l) a.act_off[l] = pl.abl_off[l] * 2;
}
// host copy of "which output slot receives the state at the end of step s"
static void ps_emit_slots(const ncde_fixed_grid_t& g, std::vector<int>* slots) {
    slots->assign((size_t)(g.n_steps > 0 ? g.n_steps : 1), -1);
    for (int64_t j = 1; j < g.n_out; ++j) (*slots)[(size_t)g.out_step[j]] = (int)j;
}


// ---------------------------------------------------------------------------------------------------------------
// Device-side loop of the adaptive solvers: a CUDA-graph WHILE node whose body is ONE attempt (every kernel of it reads its step
// size, stage times and flags from the device control block, so the body's launch parameters never change) followed by a
// one-thread kernel that sets the loop condition from ctrl->done.  The host enqueues the graph once and never looks at the
// controller again: no event, no flag copy, no stream synchronisation (north_star: "step-accept control reduced on device with
// no host sync").  NCDE_DOPRI_HOSTPOLL=1 keeps the round-1 chunked look-ahead loop for A/B runs.
// ---------------------------------------------------------------------------------------------------------------
__global__ void zero_f32_kernel(float* __restrict__ p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i < n; i += (int64_t)blockDim.x) {
        if (i >= ((int64_t)blockIdx.x + 1) * blockDim.x * 4) break;
        p[i] = 0.f;
    }
}
__global__ void adapt_cond_kernel(cudaGraphConditionalHandle handle, const AdaptCtrl* ctrl) {
    cudaGraphSetConditional(handle, ctrl->done ? 0u : 1u);
}
static bool dopri_host_poll() {
    static const bool v = getenv("NCDE_DOPRI_HOSTPOLL") != nullptr;
    return v;
}
template <typename Body>
static int graph_while_not_done(cudaStream_t st, const AdaptCtrl* ctrl, Body&& body) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaStream_t cs = nullptr;
    int rc = NCDE_OK;
    auto cleanup = [&]() {
        if (exec) cudaGraphExecDestroy(exec);     // an executable graph in flight is freed on completion
        if (graph) cudaGraphDestroy(graph);
        if (cs) cudaStreamDestroy(cs);
    };
#define NCDE_GW(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            ncde::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            cleanup();                                                                                  \
            return NCDE_ERR_CUDA;                                                                       \
        }                                                                                               \
    } while (0)
    NCDE_GW(cudaGraphCreate(&graph, 0));
    cudaGraphConditionalHandle handle;
    NCDE_GW(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    NCDE_GW(cudaGraphAddNode(&node, graph, nullptr, 0, &np));
    cudaGraph_t body_graph = np.conditional.phGraph_out[0];
    NCDE_GW(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    NCDE_GW(cudaStreamBeginCaptureToGraph(cs, body_graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    rc = body(cs);
    if (rc == NCDE_OK) adapt_cond_kernel<<<1, 1, 0, cs>>>(handle, ctrl);
    cudaGraph_t ended = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(cs, &ended);
    if (rc != NCDE_OK) { cleanup(); return rc; }
    NCDE_GW(e_end);
    NCDE_GW(cudaGraphInstantiate(&exec, graph, 0));
    NCDE_GW(cudaGraphLaunch(exec, st));
#undef NCDE_GW
    cleanup();
    return NCDE_OK;
}

}  // namespace ncde

using namespace ncde;

extern "C" const char* ncde_version(void) { return "ncde_b200 0.1.0 (sm_100a)"; }
extern "C" const char* ncde_last_error(void) { return ncde::g_error; }
extern "C" int ncde_abi_version(void) { return NCDE_ABI_VERSION; }

extern "C" int ncde_profile_enable(int class_mask) {
    ncde::g_prof_mask = class_mask;
    return NCDE_OK;
}

extern "C" int ncde_profile_read(double* ms, int64_t* count) {
    NCDE_REQUIRE(ms && count, NCDE_ERR_INVALID, "profile_read: null pointer");
    for (int c = 0; c < NCDE_PROF_CLASSES; ++c) { ms[c] = 0.0; count[c] = 0; }
    for (auto& r : ncde::g_prof) {
        NCDE_CUDA_OK(cudaEventSynchronize(r.b));
        float t = 0.f;
        NCDE_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cls] += t; count[r.cls] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    ncde::g_prof.clear();
    return NCDE_OK;
}

extern "C" size_t ncde_solve_saved_bytes(const ncde_problem_t* p, int need_grad) {
    Plan pl;
    if (!p || !need_grad || make_plan(p, &pl, true) != NCDE_OK) return 0;
    return (size_t)p->grid.n_steps * pl.n_stages * pl.stage_floats * 4 + 256;
}

static size_t adaptive_workspace_floats(const Plan& pl, int64_t n_out) {
    size_t per = 256 / 4;
    size_t n = pl.wpack_floats + per;
    n += (size_t)(2 + 7) * ((size_t)pl.H * pl.Bp + per);   // y, y1, k0..k6
    n += pl.stage_floats + per;                             // one stage of scratch activations / dX / bf16 copy
    n += sizeof(AdaptCtrl) / 4 + per + 2 * 256 + per + (size_t)n_out * 2 + per;
    return n;
}

extern "C" size_t ncde_solve_workspace_bytes(const ncde_problem_t* p, int backward) {
    Plan pl;
    if (!p || make_plan(p, &pl, p->method != NCDE_DOPRI5, false, backward != 0) != NCDE_OK) return 0;
    if (p->method == NCDE_DOPRI5) return backward ? 0 : adaptive_workspace_floats(pl, p->adaptive.n_out) * 4 + 4096;
    SwapPlan sp;
    if (backward == 2 && (pl.tc || pl.vf || pl.gated || make_swap_plan(pl, &sp) != NCDE_OK)) return 0;   // with the path gradient (fp32 path only)
    size_t fl = backward ? bwd_workspace_floats(pl, p->grid.n_steps, backward == 2 ? &sp : nullptr)
                         : fwd_workspace_floats(pl, 1) + fwd_workspace_extra_floats(pl, p->grid.n_steps, 1);
    return fl * 4 + 4096;
}

extern "C" int ncde_solve_fwd(const ncde_problem_t* p, const float* z0, float* z_out, void* saved, int need_grad,
                              void* workspace, size_t workspace_bytes, int32_t* flags, int64_t* stats,
                              int64_t* launches_out, void* stream) {
    ncde::DeviceGuard device_guard(z0);
    (void)flags;
    NCDE_REQUIRE(p && z0 && z_out, NCDE_ERR_INVALID, "solve_fwd: null pointer");
    Plan pl;
    int rc = make_plan(p, &pl, true);   // first: an unsupported problem reports why (its workspace size query returned 0)
    if (rc != NCDE_OK) return rc;
    NCDE_REQUIRE(workspace, NCDE_ERR_INVALID, "solve_fwd: null workspace");
    NCDE_REQUIRE(!need_grad || saved, NCDE_ERR_INVALID, "solve_fwd: need_grad requires a saved buffer");
    NCDE_REQUIRE(p->method != NCDE_DOPRI5, NCDE_ERR_INVALID, "solve_fwd: use ncde_solve_adaptive_fwd for dopri5");
    rc = validate_grid(p, pl);
    if (rc != NCDE_OK) return rc;
    NCDE_REQUIRE(workspace_bytes >= (fwd_workspace_floats(pl, !need_grad) +
                                     fwd_workspace_extra_floats(pl, p->grid.n_steps, !need_grad)) * 4,
                 NCDE_ERR_WORKSPACE, "solve_fwd: workspace of %zu bytes is too small", workspace_bytes);
    NCDE_REQUIRE(p->precision >= NCDE_PREC_FP32 && p->precision <= NCDE_PREC_BF16X3, NCDE_ERR_INVALID, "solve_fwd: bad precision");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t launches = 0;
    const ncde_fixed_grid_t& g = p->grid;
    const int NS = pl.n_stages;

    Carver cv{(char*)workspace, 0, workspace_bytes};
    float* wpack = cv.take(pl.wpack_floats);
    float* yT[2] = {cv.take((size_t)pl.H * pl.Bp), cv.take((size_t)pl.H * pl.Bp)};
    float* kT[NCDE_MAX_STAGES] = {};
    for (int i = 0; i < NS; ++i) kT[i] = cv.take((size_t)pl.H * pl.Bp);
    float* scratch_stage = need_grad ? nullptr : cv.take(pl.stage_floats);
    const int64_t n_st_total = g.n_steps * NS;
    float* d_stage_t = cv.take((size_t)(n_st_total > 0 ? n_st_total : 1));
    const int dx_rows = pl.vf ? pl.PC : pl.Cp;   // rows per stage of the path-only precompute
    float* dx_all = need_grad ? nullptr : cv.take((size_t)n_st_total * dx_rows * pl.Bp);
    float* e0 = pl.vf ? cv.take(4 * (size_t)pl.Bp) : nullptr;
    if (e0) {
        fill_e0_kernel<<<(unsigned)ceil_div(4 * (int64_t)pl.Bp, 256), 256, 0, st>>>(e0, pl.Bp);
        ++launches;
    }

    float* ps_ws = pl.tc_hid ? cv.take(ps_workspace_floats(pl, g.n_steps, false)) : nullptr;
    PsPlan pp;
    // the saved records of a persistent forward pass are read by the persistent backward pass only (dX/dt is kept feature-major):
    // with gradients, both passes must qualify (bf16x3 has no other path and reports what is missing)
    PsPlan pp_other;
    const bool persist = ps_eligible(p, pl, false, &pp) && (!need_grad || pl.ns == 2 || ps_eligible(p, pl, true, &pp_other));
    NCDE_REQUIRE(persist || pl.ns == 1, NCDE_ERR_UNSUPPORTED,
                 "solve_fwd: precision bf16x3 needs hidden layers of width <= 128 and every output time on a grid point");
    rc = persist ? ps_pack(p, pl, wpack, st, &launches) : pack_weights(p, pl, wpack, 0, st, &launches);
    if (rc != NCDE_OK) return rc;

    const bool use_tc = pl.tc != 0;
    if (persist) rc = NCDE_OK;
    else if (use_tc) rc = opt_in_smem(tc_field_fwd_kernel, pl.fwd_smem);
    else rc = opt_in_field(pl, false);
    if (rc != NCDE_OK) return rc;

    const dim3 tb(32, 8), tg((unsigned)ceil_div(pl.B, 32), (unsigned)ceil_div(pl.H, 32));
    to_feature_major_kernel<<<tg, tb, 0, st>>>(z0, yT[0], pl.B, pl.Bp, pl.H);
    ++launches;
    NCDE_CUDA_OK(cudaMemcpyAsync(z_out, z0, (size_t)pl.B * pl.H * 4, cudaMemcpyDeviceToDevice, st));

    HiddenFwdArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.B = pl.B; ha.Bp = pl.Bp; ha.H = pl.H; ha.C = pl.C; ha.Cp = pl.Cp; ha.R = pl.R; ha.F = pl.F; ha.Dmax = pl.Dmax;
    for (int l = 0; l <= pl.F; ++l) ha.D[l] = pl.D[l];
    for (int l = 0; l < pl.F; ++l) {
        ha.ldw[l] = pl.ldw[l]; ha.act[l] = p->mlp.act[l];
        ha.WT[l] = wpack + pl.off_WT[l]; ha.bp[l] = wpack + pl.off_bp[l];
        ha.wsm_off[l] = (int)(pl.off_WT[l] - pl.off_WT[0]);
    }
    ha.w_in_smem = pl.w_in_smem; ha.wsm_floats = (int)round_up(pl.wt_floats, 4);
    rc = opt_in_smem(hidden_fwd_kernel, pl.hid_smem_fwd);
    if (rc != NCDE_OK) return rc;
    ha.path.kind = p->path.kind; ha.path.K = (int)p->path.K; ha.path.knots = p->path.knots;
    ha.path.coeffs = p->path.coeffs; ha.path.derivs = p->path.derivs;
    ha.path.match = p->path.match; ha.path.match_terms = p->path.match_terms; ha.path.match_eps = p->path.match_eps;
    for (int i = 0; i < NS; ++i) ha.kT[i] = kT[i];

    FieldArgs fa;
    fill_field_args(fa, pl, wpack);
    TcFieldArgs ta;
    fill_tc_args(ta, pl, wpack);
    ha.KP = pl.KP;

    // dX/dt for every stage of the grid in one launch (off the sequential chain)
    if (n_st_total > 0) {
        NCDE_CUDA_OK(cudaMemcpyAsync(d_stage_t, g.stage_t, (size_t)n_st_total * 4, cudaMemcpyHostToDevice, st));
        DxAllArgs da;
        memset(&da, 0, sizeof(da));
        da.B = pl.B; da.Bp = pl.Bp; da.C = pl.C; da.Cp = pl.Cp;
        da.path = ha.path;
        da.stage_t = d_stage_t;
        da.dx_base = need_grad ? (float*)saved + pl.dx_off : dx_all;
        da.stage_stride = need_grad ? pl.stage_floats : (size_t)pl.Cp * pl.Bp;
        da.row_major = use_tc ? 1 : 0;
        if (pl.vf) {
            // X(t) or dX/dt(t) goes straight into rows H.. of the first layer's (saved) input record
            da.C = da.Cp = pl.PC;
            da.value = pl.vf == NCDE_VF_EVALUATE;
            da.dx_base = need_grad ? (float*)saved + pl.act_off[0] + (size_t)pl.H * pl.Bp : dx_all;
            da.stage_stride = need_grad ? pl.stage_floats : (size_t)pl.PC * pl.Bp;
        }
        NCDE_REQUIRE(n_st_total <= 2147483647 && ceil_div(pl.B, 32) <= 65535, NCDE_ERR_UNSUPPORTED, "solve_fwd: grid too large");
        {
            ProfScope ps(NCDE_PROF_OTHER, st);   // the path-derivative stream (bench.py reports its achieved HBM GB/s)
            dx_all_kernel<<<dim3((unsigned)n_st_total, (unsigned)ceil_div(pl.B, 32)), 256, 0, st>>>(da);
        }
        ++launches;
    }

    if (persist) {
        // ONE launch runs every stage of every step (persist_tc.cuh)
        float* const rec0p = need_grad ? (float*)saved : scratch_stage;
        const int64_t n_rec = need_grad ? n_st_total : 1;
        float* d_dt = ps_ws;
        int* d_emit = (int*)(ps_ws + round_up(g.n_steps, 64));
        int* sync = d_emit + round_up(g.n_steps, 64);
        std::vector<int> slots;
        ps_emit_slots(g, &slots);
        NCDE_CUDA_OK(cudaMemcpyAsync(d_dt, g.dt, (size_t)g.n_steps * 4, cudaMemcpyHostToDevice, st));
        NCDE_CUDA_OK(cudaMemcpyAsync(d_emit, slots.data(), (size_t)g.n_steps * 4, cudaMemcpyHostToDevice, st));
        NCDE_CUDA_OK(cudaMemsetAsync(sync, 0, (size_t)2 * pp.n_mt * sizeof(int), st));
        // feature padding (and, for the stage-input records, every column until the kernels write them) must be zero
        NCDE_CUDA_OK(cudaMemset2DAsync(rec0p + pl.abl_off[0], (need_grad ? pl.stage_floats : (size_t)pl.ns * pl.Bp * 64) * 4, 0,
                                       (size_t)pl.ns * pl.Bp * 256, (size_t)n_rec, st));
        ps_z0_record_kernel<<<(unsigned)ceil_div((int64_t)pl.B * 128, 256), 256, 0, st>>>(z0, (__nv_bfloat16*)(rec0p + pl.abl_off[0]), pl.B,
                                                                                         pl.Bp, pl.H, pl.ns);
        ++launches;
        PsMaps pm;
        rc = ps_build_maps(pl, wpack, &pm, rec0p, pl.stage_floats, n_rec, nullptr, need_grad ? (const float*)saved + pl.dx_off : dx_all,
                           need_grad ? pl.stage_floats : (size_t)pl.Cp * pl.Bp, n_st_total);
        if (rc != NCDE_OK) return rc;
        PsArgs pa;
        ps_fill_args(pa, p, pl, pp, wpack);
        pa.need_grad = need_grad ? 1 : 0;
        pa.dt = d_dt; pa.emit_idx = d_emit; pa.z_out = z_out;
        pa.yT[0] = yT[0]; pa.yT[1] = yT[1];
        for (int i = 0; i < NS; ++i) pa.kT[i] = kT[i];
        pa.rec0 = (__nv_bfloat16*)rec0p;
        pa.rec_stride = need_grad ? pl.stage_floats * 2 : 0;
        pa.dx0 = need_grad ? (const float*)saved + pl.dx_off : dx_all;
        pa.dx_stride = need_grad ? pl.stage_floats : (size_t)pl.Cp * pl.Bp;
        pa.cnt_f = sync; pa.flag_h = sync + pp.n_mt;
        static const bool trace_on = getenv("NCDE_PS_TRACE") != nullptr;   // debug: stamps of tile 0's hand-offs, dumped to stderr
        if (trace_on) {
            NCDE_CUDA_OK(cudaMalloc(&pa.trace, (kPsTraceStages * kPsTraceTiles * kPsTraceEv + 256 * 8) * 8));
            NCDE_CUDA_OK(cudaMemsetAsync(pa.trace, 0, (kPsTraceStages * kPsTraceTiles * kPsTraceEv + 256 * 8) * 8, st));
            pa.trace_g = getenv("NCDE_PS_TRACE_G") ? atoi(getenv("NCDE_PS_TRACE_G")) : 0;
        }
        const dim3 grid((unsigned)(pp.n_field + pp.n_hid));
        {
            ProfScope ps(NCDE_PROF_SOLVE_FWD, st);
            if (pl.ns == 2) {
                NCDE_CUDA_OK(cudaFuncSetAttribute(persist_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
                NCDE_CUDA_OK(launch_coop(persist_fwd_kernel<2>, grid, dim3(kPsThreads), pp.smem, st, pa, pm));
            } else {
                NCDE_CUDA_OK(cudaFuncSetAttribute(persist_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
                NCDE_CUDA_OK(launch_coop(persist_fwd_kernel<1>, grid, dim3(kPsThreads), pp.smem, st, pa, pm));
            }
        }
        ++launches;
        NCDE_CUDA_OK(cudaGetLastError());
        if (trace_on) {
            std::vector<unsigned long long> h(kPsTraceStages * kPsTraceTiles * kPsTraceEv + 256 * 8);
            NCDE_CUDA_OK(cudaStreamSynchronize(st));
            NCDE_CUDA_OK(cudaMemcpy(h.data(), pa.trace, h.size() * 8, cudaMemcpyDeviceToHost));
            cudaFree(pa.trace);
            const unsigned long long t0 = h[5];
            for (int q = 0; q < kPsTraceStages && q < g.n_steps * NS; ++q)
                for (int t = 0; t < kPsTraceTiles && t < pp.n_mt; ++t) {
                    const unsigned long long* r = h.data() + ((size_t)q * kPsTraceTiles + t) * kPsTraceEv;
                    fprintf(stderr, "pstrace fwd q=%d t=%d", q, t);
                    for (int e = 0; e < kPsTraceEv; ++e) fprintf(stderr, " %lld", r[e] ? (long long)(r[e] - t0) : -1ll);
                    fprintf(stderr, "\n");
                }
        }
        if (launches_out) *launches_out = launches;
        return NCDE_OK;
    }

    TcMapSet ms;
    if (use_tc && n_st_total > 0) {
        if (need_grad) rc = build_tc_maps(pl, wpack, &ms, (float*)saved + pl.abf_off, pl.stage_floats, n_st_total,
                                          (float*)saved + pl.dx_off, pl.stage_floats, n_st_total);
        else rc = build_tc_maps(pl, wpack, &ms, scratch_stage + pl.abf_off, 0, 1, dx_all, (size_t)pl.Cp * pl.Bp, n_st_total);
        if (rc != NCDE_OK) return rc;
    }
    // all-tensor-core path: hidden layers read / write bf16 row-major records; the stage inputs are produced by the previous
    // final-layer kernel (or by `advance` / the initial conversion)
    TcHiddenMaps hm;
    TcHiddenArgs th;
    memset(&th, 0, sizeof(th));
    float* const rec0 = need_grad ? (float*)saved : scratch_stage;
    const size_t rec_stride = need_grad ? pl.stage_floats : 0;
    auto a0_of = [&](int64_t rec) { return (__nv_bfloat16*)(rec0 + (size_t)rec * rec_stride + pl.abl_off[0]); };
    if (pl.tc_hid && n_st_total > 0) {
        rc = build_hidden_maps(pl, wpack, &hm, rec0, pl.stage_floats, need_grad ? n_st_total : 1);
        if (rc == NCDE_OK) rc = opt_in_smem(tc_hidden_fwd_kernel, tc_hid_smem_bytes());
        if (rc != NCDE_OK) return rc;
        th.B = pl.B; th.F = pl.F; th.bias = wpack + pl.off_bh;
        for (int l = 0; l < pl.F; ++l) th.act[l] = p->mlp.act[l];
        if (pl.H < 128)   // feature padding of the stage-input records must be zero, not whatever the buffer held
            NCDE_CUDA_OK(cudaMemset2DAsync(rec0 + pl.abl_off[0], (need_grad ? pl.stage_floats : (size_t)pl.Bp * 64) * 4, 0,
                                           (size_t)pl.Bp * 256, need_grad ? (size_t)n_st_total : 1, st));
        to_bf16_rows_kernel<<<(unsigned)ceil_div((int64_t)pl.B * 128, 256), 256, 0, st>>>(z0, a0_of(0), pl.B, pl.H);
        ++launches;
    }

    static const int rk4_combine[4] = {COMBINE_Y, COMBINE_RK4_S2, COMBINE_RK4_S3, COMBINE_RK4_S4};
    int cur = 0;
    int64_t j_out = 1;
    static const bool fuse_steps = getenv("NCDE_NO_STEP_FUSION") == nullptr;
    for (int64_t s = 0; s < g.n_steps; ++s) {
        const float dt = g.dt[s];
        // outputs emitted by this step; the last stage's kernel can advance the state itself when there is at most one and it
        // is the plain end-of-step state
        int64_t n_emit_step = 0;
        while (j_out + n_emit_step < g.n_out && g.out_step[j_out + n_emit_step] == s) ++n_emit_step;
        const bool fuse_adv = fuse_steps && pl.tc_hid && (n_emit_step == 0 || (n_emit_step == 1 && g.out_mode[j_out] == 1));
        for (int i = 0; i < NS; ++i) {
            float* stage = need_grad ? (float*)saved + (size_t)(s * NS + i) * pl.stage_floats : scratch_stage;
            ha.combine = p->method == NCDE_RK4_38 ? rk4_combine[i] : COMBINE_Y;
            ha.dt = dt;
            ha.yT = yT[cur];
            for (int l = 0; l <= pl.F; ++l) ha.actT[l] = stage + pl.act_off[l];
            float* dx_stage = need_grad ? stage + pl.dx_off : dx_all + (size_t)(s * NS + i) * dx_rows * pl.Bp;
            ha.dXT = nullptr;  // precomputed by dx_all_kernel
            if (pl.vf) {
                ha.n_u = pl.PC;
                ha.uT = need_grad ? stage + pl.act_off[0] + (size_t)pl.H * pl.Bp : dx_stage;
                dx_stage = e0;
            }
            ha.abf = use_tc ? (__nv_bfloat16*)(stage + pl.abf_off) : nullptr;
            ha.path.t = g.stage_t[s * NS + i];
            if (pl.tc_hid) {
                if (pl.F > 0) {
                    ProfScope ps(NCDE_PROF_HIDDEN_FWD, st);
                    th.rec = need_grad ? (int)(s * NS + i) : 0;
                    NCDE_CUDA_OK(launch_pdl(tc_hidden_fwd_kernel, dim3((unsigned)ceil_div(pl.B, kTcM)), dim3(kTcThreads), tc_hid_smem_bytes(),
                                            st, th, hm));
                    ++launches;
                }
                // this stage's final-layer kernel emits the next stage's input record
                const bool emit = i + 1 < NS;
                ta.zs_out = emit ? a0_of(need_grad ? s * NS + i + 1 : 0) : nullptr;
                ta.yT = yT[cur];
                for (int j = 0; j < NS; ++j) ta.kT[j] = kT[j];
                ta.next_combine = emit && p->method == NCDE_RK4_38 ? rk4_combine[i + 1] : COMBINE_Y;
                ta.dt = dt;
                ta.adv_ynewT = nullptr; ta.adv_ybf = nullptr; ta.adv_emit = nullptr;
                if (fuse_adv && i == NS - 1) {
                    ta.adv_ynewT = yT[cur ^ 1];
                    ta.adv_ybf = s + 1 < g.n_steps ? a0_of(need_grad ? (s + 1) * NS : 0) : nullptr;
                    ta.adv_emit = n_emit_step == 1 ? z_out + (size_t)j_out * pl.B * pl.H : nullptr;
                    ta.adv_method = p->method;
                }
            } else {
                ProfScope ps(NCDE_PROF_HIDDEN_FWD, st);
                NCDE_CUDA_OK(launch_pdl(hidden_fwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_fwd, st, ha));
                ++launches;
            }
            fa.actT = stage + pl.act_off[pl.F];
            fa.dXT = dx_stage;
            fa.koutT = kT[i];
            {
                ProfScope ps(NCDE_PROF_FIELD_FWD, st);
                if (use_tc) {
                    ta.abf = (const __nv_bfloat16*)(stage + pl.abf_off);
                    ta.dXT = dx_stage;
                    ta.koutT = kT[i];
                    ta.prefetch = 1;   // dX/dt of every stage was written by dx_all before this loop
                    { const int rc_tc = launch_tc_fwd(pl, ta, ms, st); if (rc_tc != NCDE_OK) return rc_tc; }
                    ++launches;
                } else {
                    NCDE_CUDA_OK(launch_field(pl, fa, false, st));
                    ++launches;
                }
            }
        }
        if (fuse_adv) {   // y_{n+1} was written (and emitted) by the last stage's final-layer kernel
            j_out += n_emit_step;
            cur ^= 1;
            continue;
        }
        AdvanceArgs aa;
        memset(&aa, 0, sizeof(aa));
        aa.B = pl.B; aa.Bp = pl.Bp; aa.H = pl.H; aa.method = p->method; aa.dt = dt;
        aa.yT = yT[cur]; aa.ynewT = yT[cur ^ 1];
        for (int i = 0; i < NS; ++i) aa.kT[i] = kT[i];
        aa.ybf = (pl.tc_hid && s + 1 < g.n_steps) ? a0_of(need_grad ? (s + 1) * NS : 0) : nullptr;
        bool first = true;
        while (first || (j_out < g.n_out && g.out_step[j_out] == s)) {
            aa.n_emit = 0;
            while (aa.n_emit < 4 && j_out < g.n_out && g.out_step[j_out] == s) {
                aa.emit_ptr[aa.n_emit] = z_out + (size_t)j_out * pl.B * pl.H;
                aa.emit_mode[aa.n_emit] = g.out_mode[j_out];
                aa.emit_slope[aa.n_emit] = g.out_slope[j_out];
                ++aa.n_emit; ++j_out;
            }
            NCDE_CUDA_OK(launch_pdl(advance_kernel, tg, tb, 0, st, aa));
            ++launches;
            first = false;
        }
        cur ^= 1;
    }
    NCDE_REQUIRE(j_out == g.n_out, NCDE_ERR_INVALID, "solve_fwd: %lld outputs were not covered by the grid",
                 (long long)(g.n_out - j_out));
    NCDE_CUDA_OK(cudaGetLastError());
    if (stats) {
        // fixed grid: attempted = accepted = n_steps; evaluations = n_steps * stages.  Written from the host side
        // through a stream-ordered copy of a small staging array owned by the caller is not possible here, so the
        // Python layer derives these numbers; `stats` is only used by the adaptive solver.
    }
    if (launches_out) *launches_out = launches;
    return NCDE_OK;
}

extern "C" int ncde_solve_bwd(const ncde_problem_t* p, const float* grad_out, const void* saved, float* grad_z0,
                              float* const* gW, float* const* gbias, float* grad_coeffs, void* workspace,
                              size_t workspace_bytes, int64_t* launches_out, void* stream) {
    ncde::DeviceGuard device_guard(grad_out);
    NCDE_REQUIRE(p && grad_out && saved && grad_z0 && gW && gbias && workspace, NCDE_ERR_INVALID, "solve_bwd: null pointer");
    Plan pl;
    int rc = make_plan(p, &pl, true, false, true);
    if (rc != NCDE_OK) return rc;
    rc = validate_grid(p, pl);
    if (rc != NCDE_OK) return rc;
    SwapPlan sp;
    if (grad_coeffs) {
        NCDE_REQUIRE(!pl.vf && !pl.gated, NCDE_ERR_UNSUPPORTED,
                     "solve_bwd: the gradient w.r.t. the control path is implemented for un-gated vector_field_type matmul only");
        NCDE_REQUIRE(!pl.tc, NCDE_ERR_UNSUPPORTED,
                     "solve_bwd: the gradient w.r.t. the control path is implemented for precision fp32 only");
        NCDE_REQUIRE(p->path.match == nullptr, NCDE_ERR_UNSUPPORTED,
                     "solve_bwd: the gradient w.r.t. the control path is not implemented for gradient-matched (smoothed) paths");
        rc = make_swap_plan(pl, &sp);
        if (rc != NCDE_OK) return rc;
    }
    NCDE_REQUIRE(workspace_bytes >= bwd_workspace_floats(pl, p->grid.n_steps, grad_coeffs ? &sp : nullptr) * 4, NCDE_ERR_WORKSPACE,
                 "solve_bwd: workspace of %zu bytes is too small", workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t launches = 0;
    const ncde_fixed_grid_t& g = p->grid;
    const ncde_mlp_t& m = p->mlp;
    const int NS = pl.n_stages;
    const size_t nHB = (size_t)pl.H * pl.Bp;

    Carver cv{(char*)workspace, 0, workspace_bytes};
    float* wpack = cv.take(pl.wpack_floats);
    const size_t nHBp = round_up(pl.H, 4) * (size_t)pl.Bp;   // rows H..Hp-1: zero padding read by the path-gradient launch
    float* gyT = cv.take(nHBp);
    float* gkT[NCDE_MAX_STAGES] = {};
    for (int i = 0; i < NS; ++i) gkT[i] = cv.take(nHBp);
    float* e0 = pl.vf ? cv.take(4 * (size_t)pl.Bp) : nullptr;
    float* W3Ts = nullptr;
    float* b3ps = nullptr;
    float* gdXT[NCDE_MAX_STAGES] = {};
    if (grad_coeffs) {
        W3Ts = cv.take(sp.w3t_floats);
        b3ps = cv.take(sp.b3p_floats);
        for (int i = 0; i < NS; ++i) gdXT[i] = cv.take((size_t)pl.C * pl.Bp);
    }
    float* P = cv.take((size_t)pl.n_hg * pl.Bp * pl.DFP);
    float* dpreT[NCDE_MAX_STAGES][NCDE_MAX_LAYERS] = {};
    for (int i = 0; i < NS; ++i)
        for (int l = 0; l < pl.F; ++l) dpreT[i][l] = cv.take((size_t)pl.Dp4[l + 1] * pl.Bp);
    float* dW3acc = cv.take((size_t)pl.n_bt * pl.Np * pl.DFP);
    float* db3acc = cv.take((size_t)pl.n_bt * pl.Np);
    // all-tensor-core path: one bf16 record for dL/d(pre-activation of the last hidden layer), padded fp32 accumulators
    const bool tc_hid = pl.tc_hid && pl.F > 0;
    const int64_t n_rec_all = g.n_steps * NS;
    float* dpre_rec = tc_hid ? cv.take((size_t)(n_rec_all > 0 ? n_rec_all : 1) * pl.F * pl.ns * pl.Bp * 64) : nullptr;   // [rec][layer][part][Bp][128] bf16
    int* tile_count = tc_hid ? (int*)cv.take((size_t)(pl.Bp / 128)) : nullptr;   // fused reduction: column blocks done per batch tile
    float* dWh_acc = tc_hid ? cv.take((size_t)pl.F * 128 * 128) : nullptr;
    float* dbh_acc = tc_hid ? cv.take((size_t)pl.F * 128) : nullptr;

    float* ps_ws = pl.tc_hid ? cv.take(ps_workspace_floats(pl, g.n_steps, true)) : nullptr;
    PsPlan pp;
    PsPlan pp_other;
    const bool persist = !grad_coeffs && ps_eligible(p, pl, true, &pp) && (pl.ns == 2 || ps_eligible(p, pl, false, &pp_other));
    NCDE_REQUIRE(persist || pl.ns == 1, NCDE_ERR_UNSUPPORTED,
                 "solve_bwd: precision bf16x3 needs hidden layers of width <= 128 and every output time on a grid point");
    if (persist) {
        // ONE launch runs the backward pass of every stage (persist_tc.cuh), then the hidden weight gradients in one split-K launch
        const int64_t n_rec = n_rec_all;
        float* d_dt = ps_ws;
        int* d_emit = (int*)(ps_ws + round_up(g.n_steps, 64));
        int* sync = d_emit + round_up(g.n_steps, 64);
        float* dAT = (float*)(sync + round_up(2 * pp.n_mt, 64));
        float* gy2 = dAT + (size_t)pp.n_mt * 128 * 128 + 64;
        std::vector<int> slots;
        ps_emit_slots(g, &slots);
        NCDE_CUDA_OK(cudaMemcpyAsync(d_dt, g.dt, (size_t)g.n_steps * 4, cudaMemcpyHostToDevice, st));
        NCDE_CUDA_OK(cudaMemcpyAsync(d_emit, slots.data(), (size_t)g.n_steps * 4, cudaMemcpyHostToDevice, st));
        NCDE_CUDA_OK(cudaMemsetAsync(sync, 0, (size_t)2 * pp.n_mt * sizeof(int), st));
        NCDE_CUDA_OK(cudaMemsetAsync(dAT, 0, (size_t)pp.n_mt * 128 * 128 * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(gyT, 0, nHB * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(gy2, 0, nHB * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(dW3acc, 0, (size_t)pp.n_part * pl.Np * 128 * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(db3acc, 0, (size_t)pp.n_part * pl.Np * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(dWh_acc, 0, (size_t)pl.F * (128 * 128) * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(dbh_acc, 0, (size_t)pl.F * 128 * 4, st));
        rc = ps_pack(p, pl, wpack, st, &launches);
        if (rc != NCDE_OK) return rc;
        PsMaps pm;
        rc = ps_build_maps(pl, wpack, &pm, (const float*)saved, pl.stage_floats, n_rec, dpre_rec, (const float*)saved + pl.dx_off,
                           pl.stage_floats, n_rec, ps_bwd_narrow(pl.Hg) ? kPsXPitchN : kPsXPitch);
        if (rc != NCDE_OK) return rc;
        PsArgs pa;
        ps_fill_args(pa, p, pl, pp, wpack);
        pa.need_grad = 1;
        pa.dt = d_dt; pa.emit_idx = d_emit; pa.grad_out = grad_out;
        pa.yT[0] = gyT; pa.yT[1] = gy2;
        for (int i = 0; i < NS; ++i) pa.kT[i] = gkT[i];
        pa.rec0 = (__nv_bfloat16*)const_cast<void*>(saved);
        pa.rec_stride = pl.stage_floats * 2;
        pa.dx0 = (const float*)saved + pl.dx_off;
        pa.dx_stride = pl.stage_floats;
        pa.dAT = dAT; pa.dW3acc = dW3acc; pa.db3acc = db3acc; pa.dpre0 = (__nv_bfloat16*)dpre_rec;
        pa.cnt_f = sync; pa.flag_h = sync + pp.n_mt;
        static const bool trace_on = getenv("NCDE_PS_TRACE") != nullptr;
        if (trace_on) {
            NCDE_CUDA_OK(cudaMalloc(&pa.trace, (kPsTraceStages * kPsTraceTiles * kPsTraceEv + 256 * 8) * 8));
            NCDE_CUDA_OK(cudaMemsetAsync(pa.trace, 0, (kPsTraceStages * kPsTraceTiles * kPsTraceEv + 256 * 8) * 8, st));
            pa.trace_g = getenv("NCDE_PS_TRACE_G") ? atoi(getenv("NCDE_PS_TRACE_G")) : 0;
        }
        const dim3 grid((unsigned)(pp.n_field + pp.n_hid));
        {
            ProfScope ps(NCDE_PROF_SOLVE_BWD, st);
            if (pl.ns == 2) {
                NCDE_CUDA_OK(cudaFuncSetAttribute(persist_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
                NCDE_CUDA_OK(launch_coop(persist_bwd_kernel<2>, grid, dim3(kPsThreads), pp.smem, st, pa, pm));
            } else {
                NCDE_CUDA_OK(cudaFuncSetAttribute(persist_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.smem));
                NCDE_CUDA_OK(launch_coop(persist_bwd_kernel<1>, grid, dim3(kPsThreads), pp.smem, st, pa, pm));
            }
        }
        ++launches;
        if (trace_on) {
            std::vector<unsigned long long> h(kPsTraceStages * kPsTraceTiles * kPsTraceEv + 256 * 8);
            NCDE_CUDA_OK(cudaStreamSynchronize(st));
            NCDE_CUDA_OK(cudaMemcpy(h.data(), pa.trace, h.size() * 8, cudaMemcpyDeviceToHost));
            cudaFree(pa.trace);
            const unsigned long long t0 = h[1];
            for (int q = 0; q < kPsTraceStages && q < g.n_steps * NS; ++q)
                for (int t = 0; t < kPsTraceTiles && t < pp.n_mt; ++t) {
                    const unsigned long long* r = h.data() + ((size_t)q * kPsTraceTiles + t) * kPsTraceEv;
                    fprintf(stderr, "pstrace bwd q=%d t=%d", q, t);
                    for (int e = 0; e < kPsTraceEv; ++e) fprintf(stderr, " %lld", r[e] ? (long long)(r[e] - t0) : -1ll);
                    fprintf(stderr, "\n");
                }
            for (int gg = 0; gg < 256; ++gg) {
                const unsigned long long* r = h.data() + (size_t)kPsTraceStages * kPsTraceTiles * kPsTraceEv + gg * 8;
                if (!r[1]) continue;
                fprintf(stderr, "pstrace_all bwd g=%d", gg);
                for (int e = 0; e < 7; ++e) fprintf(stderr, " %lld", r[e] ? (long long)(r[e] - t0) : -1ll);
                fprintf(stderr, " %lld\n", (long long)r[7] - 1);
            }
        }
        ps_gy_final_kernel<<<(unsigned)ceil_div((int64_t)nHB, 256), 256, 0, st>>>(gyT, gkT[0], gkT[1], gkT[2], gkT[3], NS, (int64_t)nHB);
        ++launches;
        {
            PsWgradArgs wga;
            memset(&wga, 0, sizeof(wga));
            wga.F = pl.F; wga.n_rec = (int)n_rec; wga.n_mt = pp.n_mt;
            const int64_t units = n_rec * wga.n_mt;
            const int n_split = device_sm_count() / pl.F;
            wga.n_split = (int)(n_split > units ? units : n_split);
            for (int l = 0; l < pl.F; ++l) {
                wga.dWacc[l] = dWh_acc + (size_t)pl.first_of_slot[l] * 128 * 128;
                wga.dbacc[l] = dbh_acc + (size_t)pl.first_of_slot[l] * 128;
            }
            ProfScope ps(NCDE_PROF_HIDDEN_WGRAD, st);
            if (pl.ns == 2) {
                NCDE_CUDA_OK(cudaFuncSetAttribute(ps_hidden_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_wgrad_smem_bytes(2)));
                ps_hidden_wgrad_kernel<2><<<dim3(pl.F, wga.n_split), kTcThreads, ps_wgrad_smem_bytes(2), st>>>(wga, pm);
            } else {
                NCDE_CUDA_OK(cudaFuncSetAttribute(ps_hidden_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_wgrad_smem_bytes(1)));
                ps_hidden_wgrad_kernel<1><<<dim3(pl.F, wga.n_split), kTcThreads, ps_wgrad_smem_bytes(1), st>>>(wga, pm);
            }
        }
        ++launches;
        for (int l = 0; l < pl.F; ++l) {
            if (pl.first_of_slot[l] != l) continue;
            const int n = m.out_dim[l] * m.in_dim[l];
            unpack_hidden_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dWh_acc + (size_t)l * 128 * 128, dbh_acc + (size_t)l * 128,
                                                                                gW[l], gbias[l], m.out_dim[l], m.in_dim[l]);
            ++launches;
        }
        const dim3 tbp(32, 8), tgp((unsigned)ceil_div(pl.B, 32), (unsigned)ceil_div(pl.H, 32));
        from_feature_major_kernel<<<tgp, tbp, 0, st>>>(gyT, grad_out, grad_z0, pl.B, pl.Bp, pl.H);
        ++launches;
        NCDE_REQUIRE(gW[pl.F] != nullptr, NCDE_ERR_INVALID, "solve_bwd: gW of the final layer is null");
        {
            const int64_t n = (int64_t)pl.H * pl.C * pl.DF;
            unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dW3acc, db3acc, gW[pl.F], gbias[pl.F], pl.H, pl.C, pl.Cp,
                                                                                 pl.Hg, pl.Npad, pl.DF, 128, pl.Np, pp.n_part);
            ++launches;
        }
        NCDE_CUDA_OK(cudaGetLastError());
        if (launches_out) *launches_out = launches;
        return NCDE_OK;
    }
    rc = pack_weights(p, pl, wpack, 1, st, &launches);
    if (rc != NCDE_OK) return rc;
    const bool use_tc = pl.tc != 0;
    if (use_tc) rc = (pl.bwd_ew == 16 ? opt_in_smem(tc_field_bwd_kernel<16>, pl.bwd_smem) : opt_in_smem(tc_field_bwd_kernel<8>, pl.bwd_smem));
    else rc = opt_in_field(pl, true);
    if (rc != NCDE_OK) return rc;
    TcMapSet ms;
    if (use_tc && g.n_steps > 0) {
        rc = build_tc_maps(pl, wpack, &ms, (const float*)saved + pl.abf_off, pl.stage_floats, g.n_steps * NS,
                           (const float*)saved + pl.dx_off, pl.stage_floats, g.n_steps * NS);
        if (rc != NCDE_OK) return rc;
    }

    if (e0) {
        fill_e0_kernel<<<(unsigned)ceil_div(4 * (int64_t)pl.Bp, 256), 256, 0, st>>>(e0, pl.Bp);
        ++launches;
    }
    NCDE_CUDA_OK(cudaMemsetAsync(gyT, 0, nHB * 4, st));
    NCDE_CUDA_OK(cudaMemsetAsync(dW3acc, 0, (size_t)pl.n_bt * pl.Np * pl.DFP * 4, st));
    NCDE_CUDA_OK(cudaMemsetAsync(db3acc, 0, (size_t)pl.n_bt * pl.Np * 4, st));

    FieldArgs fa;
    fill_field_args(fa, pl, wpack);
    fa.P = P; fa.dW3acc = dW3acc; fa.db3acc = db3acc;
    TcFieldArgs ta;
    fill_tc_args(ta, pl, wpack);
    ta.P = P; ta.dW3acc = dW3acc; ta.db3acc = db3acc;
    // path gradient: the forward final-layer kernel on the (h <-> c)-exchanged problem, one launch per stage
    FieldArgs fs;
    PathGradArgs pg;
    memset(&fs, 0, sizeof(fs));
    memset(&pg, 0, sizeof(pg));
    if (grad_coeffs) {
        const int64_t n = (int64_t)sp.Np * pl.DF;
        pack_final_swapped_kernel<<<(unsigned)ceil_div(n > sp.Np ? n : sp.Np, 256), 256, 0, st>>>(
            m.W[pl.F], m.bias[pl.F], W3Ts, b3ps, pl.H, pl.C, sp.Hp, pl.DF, sp.Np);
        ++launches;
        if (sp.Hp > pl.H)
            for (int i = 0; i < NS; ++i)
                NCDE_CUDA_OK(cudaMemsetAsync(gkT[i] + nHB, 0, (nHBp - nHB) * 4, st));
        fs.B = pl.B; fs.Bp = pl.Bp; fs.H = pl.C; fs.Cp = sp.Hp; fs.DF = pl.DF; fs.DFP = pl.DFP; fs.S = sp.S; fs.Hg = sp.Hg;
        fs.n_hg = sp.n_hg; fs.Np = sp.Np; fs.Bt = sp.Bt;
        fs.W3T = W3Ts; fs.b3p = b3ps;
        rc = sp.TM == 8 ? opt_in_smem(field_fwd_kernel<8>, sp.smem) : opt_in_smem(field_fwd_kernel<4>, sp.smem);
        if (rc != NCDE_OK) return rc;
        pg.B = pl.B; pg.Bp = pl.Bp; pg.C = pl.C; pg.n_stage = NS; pg.kind = p->path.kind; pg.K = (int)p->path.K;
        pg.knots = p->path.knots; pg.grad_coeffs = grad_coeffs;
        for (int i = 0; i < NS; ++i) pg.gdXT[i] = gdXT[i];
    }

    HiddenBwdArgs hb;
    memset(&hb, 0, sizeof(hb));
    hb.B = pl.B; hb.Bp = pl.Bp; hb.H = pl.H; hb.R = pl.R; hb.F = pl.F; hb.Dmax = pl.Dmax; hb.DFP = pl.DFP; hb.n_hg = pl.n_hg;
    for (int l = 0; l <= pl.F; ++l) hb.D[l] = pl.D[l];
    for (int l = 0; l < pl.F; ++l) {
        hb.act[l] = m.act[l]; hb.W[l] = wpack + pl.off_WR[l]; hb.ldi[l] = pl.ldi[l];
        hb.wsm_off[l] = (int)(pl.off_WR[l] - pl.off_WR[0]);
    }
    hb.w_in_smem = pl.w_in_smem; hb.wsm_floats = (int)round_up(pl.wr_floats, 4);
    hb.P = P; hb.gyT = gyT;
    TcHiddenBwdMaps hbm;
    TcHiddenBwdArgs thb;
    PReduceArgs pr;
    memset(&thb, 0, sizeof(thb));
    memset(&pr, 0, sizeof(pr));
    if (tc_hid) {
        NCDE_REQUIRE(pl.F <= 3, NCDE_ERR_UNSUPPORTED, "solve_bwd: at most 3 hidden layers on the tensor-core path");
        memset(&hbm, 0, sizeof(hbm));
        const int64_t n_rec = g.n_steps * NS;
        rc = make_map(&hbm.W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpack + pl.off_Wh, 128, 128, (uint64_t)pl.F, 256, 128 * 256, 64, 128,
                      CU_TENSOR_MAP_SWIZZLE_128B);
        for (int l = 0; l < pl.F && rc == NCDE_OK; ++l)
            rc = make_map(&hbm.act[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const float*)saved + pl.abl_off[l], 128, (uint64_t)pl.B,
                          (uint64_t)(n_rec < 1 ? 1 : n_rec), 256, n_rec > 1 ? pl.stage_floats * 4 : 0, 64, kTcM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == NCDE_OK)
            rc = make_map(&hbm.dpre, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dpre_rec, 128, (uint64_t)pl.B,
                          (uint64_t)(n_rec < 1 ? 1 : n_rec) * pl.F, 256, (size_t)pl.Bp * 256, 64, kTcM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == NCDE_OK) rc = opt_in_smem(tc_hidden_bwd_kernel, tc_hid_bwd_smem_bytes());
        if (rc == NCDE_OK) rc = opt_in_smem(tc_hidden_wgrad_kernel, tc_hid_wgrad_smem_bytes());
        if (rc != NCDE_OK) return rc;
        NCDE_CUDA_OK(cudaMemsetAsync(tile_count, 0, (size_t)(pl.Bp / 128) * sizeof(int), st));
        NCDE_CUDA_OK(cudaMemsetAsync(dWh_acc, 0, (size_t)pl.F * (128 * 128) * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(dbh_acc, 0, (size_t)pl.F * 128 * 4, st));
        thb.B = pl.B; thb.Bp = pl.Bp; thb.H = pl.H; thb.F = pl.F;
        for (int l = 0; l < pl.F; ++l) thb.act[l] = m.act[l];
        pr.B = pl.B; pr.Bp = pl.Bp; pr.n_hg = pl.n_hg; pr.DFP = pl.DFP; pr.act = m.act[pl.F - 1]; pr.P = P;
    }
    rc = opt_in_smem(hidden_bwd_kernel, pl.hid_smem_bwd);
    if (rc == NCDE_OK) rc = opt_in_smem(hidden_wgrad_kernel, 36 * 1024);
    if (rc != NCDE_OK) return rc;

    // weight-gradient tiles grouped by slot
    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.B = pl.B; wa.Bp = pl.Bp; wa.n_split = pl.wg_split; wa.rows_per_split = pl.wg_rows;
    int total_tiles = 0;
    for (int l = 0; l < pl.F; ++l) {
        int sidx = -1;
        for (int s2 = 0; s2 < wa.n_slots; ++s2)
            if (wa.gW[s2] == gW[l]) sidx = s2;
        if (sidx < 0) {
            sidx = wa.n_slots++;
            wa.gW[sidx] = gW[l]; wa.gb[sidx] = gbias[l];
            wa.Dout[sidx] = m.out_dim[l]; wa.Din[sidx] = m.in_dim[l];
            NCDE_REQUIRE(gW[l] != nullptr, NCDE_ERR_INVALID, "solve_bwd: gW[%d] is null", l);
        } else {
            NCDE_REQUIRE(wa.Dout[sidx] == m.out_dim[l] && wa.Din[sidx] == m.in_dim[l] && m.slot[l] == m.slot[wa.lay[sidx][0]],
                         NCDE_ERR_INVALID, "solve_bwd: layers sharing a gradient buffer must share a slot and shape");
        }
        wa.lay[sidx][wa.n_lay[sidx]++] = l;
    }
    for (int s2 = 0; s2 < wa.n_slots; ++s2) {
        wa.tile_begin[s2] = total_tiles;
        total_tiles += (int)(ceil_div(wa.Dout[s2], kWgTile) * ceil_div(wa.Din[s2], kWgTile));
        const size_t nW = (size_t)pl.wg_split * wa.Dout[s2] * wa.Din[s2], nb = (size_t)pl.wg_split * wa.Dout[s2];
        wa.gWp[s2] = cv.take(nW);
        wa.gbp[s2] = cv.take(nb);
        NCDE_CUDA_OK(cudaMemsetAsync(wa.gWp[s2], 0, nW * 4, st));
        NCDE_CUDA_OK(cudaMemsetAsync(wa.gbp[s2], 0, nb * 4, st));
    }
    wa.tile_begin[wa.n_slots] = total_tiles;
    NCDE_REQUIRE(cv.used <= workspace_bytes, NCDE_ERR_WORKSPACE, "solve_bwd: workspace accounting error");
    NCDE_REQUIRE(gW[pl.F] != nullptr, NCDE_ERR_INVALID, "solve_bwd: gW of the final layer is null");

    const dim3 tb(32, 8), tg((unsigned)ceil_div(pl.B, 32), (unsigned)ceil_div(pl.H, 32));
    const unsigned ew_grid = (unsigned)ceil_div((int64_t)nHB, 256);
    const float third = 0.3333333432674408f;

    int stages_done = 0;     // fused reduction: target of the per-tile counters
    static const bool fuse_steps = getenv("NCDE_NO_STEP_FUSION") == nullptr;
    bool pre_injected = false;   // the output gradients that enter before this step's stages were already added by the previous step's last kernel
    int64_t j_hi = g.n_out;  // outputs [j_lo, j_hi) belong to the current step
    for (int64_t s = g.n_steps - 1; s >= 0; --s) {
        int64_t j_lo = j_hi;
        while (j_lo - 1 >= 1 && g.out_step[j_lo - 1] == s) --j_lo;
        const float dt = g.dt[s];
        // All-tensor-core path: the last backward stage of the step (stage 0) writes gy <- gy + sum_i dz_i itself, and — when this
        // step injects nothing afterwards and the next (earlier) step injects at most one output gradient beforehand — that one too.
        // (measured: folding this elementwise work into the 8-CTA chain kernel costs 18 us per step against 8 us for the two wide
        //  elementwise launches it replaces, so it is off unless NCDE_FOLD_GY is set)
        static const bool fold_gy_enabled = getenv("NCDE_FOLD_GY") != nullptr;
        bool fold_gy = fuse_steps && tc_hid && fold_gy_enabled;
        const float* fold_gout = nullptr;
        float fold_scale = 0.f;
        bool fold_next_pre = false;
        if (fold_gy) {
            bool after_any = false;
            for (int64_t j = j_lo; j < j_hi; ++j)
                if (g.out_mode[j] == 0 || (g.out_mode[j] == 2 && g.out_slope[j] != 1.f)) after_any = true;
            if (!after_any && s > 0) {
                int64_t k_lo = j_lo;
                while (k_lo - 1 >= 1 && g.out_step[k_lo - 1] == s - 1) --k_lo;
                int n_pre = 0;
                for (int64_t j = k_lo; j < j_lo; ++j) {
                    const float sc = g.out_mode[j] == 1 ? 1.f : (g.out_mode[j] == 2 ? g.out_slope[j] : 0.f);
                    if (sc != 0.f) { ++n_pre; fold_gout = grad_out + (size_t)j * pl.B * pl.H; fold_scale = sc; }
                }
                if (n_pre <= 1) fold_next_pre = true;
                else { fold_gout = nullptr; fold_scale = 0.f; }
            }
        }
        for (int64_t j = j_lo; j < j_hi && !pre_injected; ++j) {
            const float sc = g.out_mode[j] == 1 ? 1.f : (g.out_mode[j] == 2 ? g.out_slope[j] : 0.f);
            if (sc != 0.f) {
                NCDE_CUDA_OK(launch_pdl(add_out_grad_kernel, tg, tb, 0, st, gyT, grad_out + (size_t)j * pl.B * pl.H, sc, pl.B, pl.Bp, pl.H));
                ++launches;
            }
        }
        if (!tc_hid) {
            NCDE_CUDA_OK(launch_pdl(rk_bwd_begin_kernel, dim3(ew_grid), dim3(256), 0, st, (const float*)gyT, gkT[0], gkT[1], gkT[2], gkT[3], (int)p->method, dt, (int64_t)nHB));
            ++launches;
        }
        for (int i = NS - 1; i >= 0; --i) {
            const float* stage = (const float*)saved + (size_t)(s * NS + i) * pl.stage_floats;
            fa.actT = stage + pl.act_off[pl.F];
            fa.dXT = pl.vf ? e0 : stage + pl.dx_off;
            fa.gkT = gkT[i];
            {
                ProfScope ps(NCDE_PROF_FIELD_BWD, st);
                if (use_tc) {
                    ta.abf = (const __nv_bfloat16*)(stage + pl.abf_off);
                    ta.dXT = stage + pl.dx_off;
                    ta.gkT = gkT[i];
                    ta.prefetch = 1;   // activations and dX/dt are records saved by the forward pass
                    if (tc_hid) {
                        // dL/dk_i = c_i dt gy1 + sum over later stages q of d(stage input q)/dk_i * dz_q; the dz_q live in gkT[q]
                        ta.gy1T = gyT;
                        ta.p_transposed = 1;
                        {
                            static const bool gk_early = getenv("NCDE_GK_EARLY") != nullptr;
                            ta.gk_early = gk_early ? 1 : 0;
                        }
                        ta.n_dz = 0;
                        if (p->method == NCDE_RK4_38) {
                            ta.gcoef = (i == 0 || i == 3) ? dt * 0.125f : 3.f * (dt * 0.125f);
                            const float c[4][4] = {{0, dt * third, -(dt * third), dt}, {0, 0, dt, -dt}, {0, 0, 0, dt}, {0, 0, 0, 0}};
                            for (int q = i + 1; q < NS; ++q)
                                if (c[i][q] != 0.f) { ta.dzT[ta.n_dz] = gkT[q]; ta.dzcoef[ta.n_dz] = c[i][q]; ++ta.n_dz; }
                        } else {
                            ta.gcoef = dt;
                        }
                    }
                    { const int rc_tc = launch_tc_bwd(pl, ta, ms, st); if (rc_tc != NCDE_OK) return rc_tc; }
                    ++launches;
                } else {
                    NCDE_CUDA_OK(launch_field(pl, fa, true, st));
                    ++launches;
                }
            }
            if (grad_coeffs) {
                // dL/d(dX/dt)[b,c] = sum_h tanh(pre[b,h,c]) gk_i[b,h]
                fs.actT = stage + pl.act_off[pl.F];
                fs.dXT = gkT[i];
                fs.koutT = gdXT[i];
                const dim3 sg(sp.n_hg, sp.n_bt);
                if (sp.TM == 8) NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<8>, sg, dim3(kThreads), sp.smem, st, fs));
                else NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<4>, sg, dim3(kThreads), sp.smem, st, fs));
                ++launches;
            }
            for (int l = 0; l <= pl.F; ++l) hb.actT[l] = stage + pl.act_off[l];
            for (int l = 0; l < pl.F; ++l) hb.dpreT[l] = dpreT[i][l];
            // d(stage input)/d(k_j): rk_common.py:111-113
            hb.n_k = 0;
            for (int j = 0; j < NCDE_MAX_STAGES; ++j) { hb.gkT[j] = nullptr; hb.kcoef[j] = 0.f; }
            if (p->method == NCDE_RK4_38) {
                if (i == 1) { hb.n_k = 1; hb.kcoef[0] = dt * third; }
                if (i == 2) { hb.n_k = 2; hb.kcoef[0] = -(dt * third); hb.kcoef[1] = dt; }
                if (i == 3) { hb.n_k = 3; hb.kcoef[0] = dt; hb.kcoef[1] = -dt; hb.kcoef[2] = dt; }
                for (int j = 0; j < hb.n_k; ++j) hb.gkT[j] = gkT[j];
            }
            if (tc_hid) {
                ProfScope ps(NCDE_PROF_HIDDEN_BWD, st);
                pr.aF = (const __nv_bfloat16*)(stage + pl.abl_off[pl.F]);
                pr.dpre = (__nv_bfloat16*)dpre_rec + ((size_t)(s * NS + i) * pl.F + (pl.F - 1)) * pl.Bp * 128;
                thb.rec = (int)(s * NS + i);
                thb.dz_out = gkT[i];   // the stage-input gradient of stage i takes the place of the (unused) gk_i array
                thb.gy_io = nullptr; thb.gout = nullptr; thb.n_other = 0;
                if (fold_gy && i == 0) {
                    thb.gy_io = gyT;
                    for (int q = 1; q < NS; ++q) thb.dz_other[thb.n_other++] = gkT[q];
                    thb.gout = fold_next_pre ? fold_gout : nullptr;
                    thb.gout_scale = fold_scale;
                }
                static const bool fused_reduce = getenv("NCDE_NO_FUSED_PREDUCE") == nullptr;
                if (fused_reduce) {
                    // one launch: 16 CTAs per batch tile reduce the partials, the first of them then runs the GEMM chain
                    thb.pr = pr;
                    thb.tile_count = tile_count;
                    thb.count_target = 16 * (++stages_done);
                    NCDE_CUDA_OK(launch_pdl(tc_hidden_bwd_kernel, dim3((unsigned)ceil_div(pl.B, kTcM), 16), dim3(kTcThreads),
                                            tc_hid_bwd_smem_bytes(), st, thb, hbm));
                    launches += 1;
                } else {
                    NCDE_CUDA_OK(launch_pdl(p_reduce_kernel, dim3((unsigned)ceil_div(pl.B, 128), 16), dim3(128), 0, st, pr));
                    thb.pr.P = nullptr;
                    NCDE_CUDA_OK(launch_pdl(tc_hidden_bwd_kernel, dim3((unsigned)ceil_div(pl.B, kTcM)), dim3(kTcThreads),
                                            tc_hid_bwd_smem_bytes(), st, thb, hbm));
                    launches += 2;
                }
                continue;
            }
            {
                ProfScope ps(NCDE_PROF_HIDDEN_BWD, st);
                NCDE_CUDA_OK(launch_pdl(hidden_bwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_bwd, st, hb));
            }
            ++launches;
        }
        if (grad_coeffs) {
            for (int i = 0; i < NS; ++i) pg.t[i] = g.stage_t[s * NS + i];
            NCDE_CUDA_OK(launch_pdl(path_grad_kernel, dim3((unsigned)ceil_div(pl.B, 32)), dim3(256), 0, st, pg));
            ++launches;
        }
        if (tc_hid && !fold_gy) {
            NCDE_CUDA_OK(launch_pdl(gy_accumulate_kernel, dim3(ew_grid), dim3(256), 0, st, gyT, (const float*)gkT[0], (const float*)gkT[1],
                                    (const float*)gkT[2], (const float*)gkT[3], NS, (int64_t)nHB));
            ++launches;
        }
        pre_injected = fold_gy && fold_next_pre;
        if (pl.F > 0 && !tc_hid) {
            // hidden weight gradients of all stages of this step in one launch (off the sequential chain)
            wa.n_stage = NS;
            for (int i = 0; i < NS; ++i) {
                const float* stage = (const float*)saved + (size_t)(s * NS + i) * pl.stage_floats;
                for (int l = 0; l < pl.F; ++l) { wa.actT[i][l] = stage + pl.act_off[l]; wa.dpreT[i][l] = dpreT[i][l]; }
            }
            {
                ProfScope ps(NCDE_PROF_HIDDEN_WGRAD, st);
                NCDE_CUDA_OK(launch_pdl(hidden_wgrad_kernel, dim3(total_tiles, pl.wg_split), dim3(kThreads), 0, st, wa));
            }
            ++launches;
        }
        for (int64_t j = j_lo; j < j_hi; ++j) {
            const float sc = g.out_mode[j] == 0 ? 1.f : (g.out_mode[j] == 2 ? 1.f - g.out_slope[j] : 0.f);
            if (sc != 0.f) {
                NCDE_CUDA_OK(launch_pdl(add_out_grad_kernel, tg, tb, 0, st, gyT, grad_out + (size_t)j * pl.B * pl.H, sc, pl.B, pl.Bp, pl.H));
                ++launches;
            }
        }
        j_hi = j_lo;
    }
    if (tc_hid && n_rec_all > 0) {
        // weight / bias gradients of the hidden layers for the whole pass: one split-K launch over every (stage, batch tile)
        TcHiddenWgradMaps wm;
        TcHiddenWgradArgs wga;
        memset(&wm, 0, sizeof(wm));
        memset(&wga, 0, sizeof(wga));
        for (int l = 0; l < pl.F; ++l) wm.act[l] = hbm.act[l];
        wm.dpre = hbm.dpre;
        wga.F = pl.F; wga.n_rec = (int)n_rec_all; wga.n_mt = (int)ceil_div(pl.B, kTcM);
        const int64_t units = n_rec_all * wga.n_mt;
        int n_split = kNumSMs / pl.F;
        wga.n_split = (int)(n_split > units ? units : n_split);
        for (int l = 0; l < pl.F; ++l) {
            // layers that share a parameter slot add into the accumulator of the first of them
            wga.dWacc[l] = dWh_acc + (size_t)pl.first_of_slot[l] * 128 * 128;
            wga.dbacc[l] = dbh_acc + (size_t)pl.first_of_slot[l] * 128;
        }
        {
            ProfScope ps(NCDE_PROF_HIDDEN_WGRAD, st);
            NCDE_CUDA_OK(launch_pdl(tc_hidden_wgrad_kernel, dim3(pl.F, wga.n_split), dim3(kTcThreads), tc_hid_wgrad_smem_bytes(), st, wga, wm));
        }
        ++launches;
    }
    if (tc_hid) {
        for (int l = 0; l < pl.F; ++l) {
            if (pl.first_of_slot[l] != l) continue;   // the slot's accumulator already holds the sum over its layers
            const int n = m.out_dim[l] * m.in_dim[l];
            unpack_hidden_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dWh_acc + (size_t)l * 128 * 128, dbh_acc + (size_t)l * 128,
                                                                                gW[l], gbias[l], m.out_dim[l], m.in_dim[l]);
            ++launches;
        }
    } else if (pl.F > 0) {
        int nmax = 0;
        for (int s2 = 0; s2 < wa.n_slots; ++s2) nmax = wa.Dout[s2] * wa.Din[s2] > nmax ? wa.Dout[s2] * wa.Din[s2] : nmax;
        hidden_wgrad_reduce_kernel<<<dim3((unsigned)ceil_div(nmax, 256), wa.n_slots), 256, 0, st>>>(wa);
        ++launches;
    }
    // solution[0] = y0 (solvers.py:95): its gradient flows straight to z0
    from_feature_major_kernel<<<tg, tb, 0, st>>>(gyT, grad_out, grad_z0, pl.B, pl.Bp, pl.H);
    ++launches;
    {
        const int64_t n = (int64_t)pl.H * pl.C * pl.DF;
        if (pl.gated) {
            // interleaved columns: 2c = sigmoid head (gradients go to gW[n_layers] / gbias[n_layers]), 2c + 1 = tanh head
            NCDE_REQUIRE(gW[m.n_layers] != nullptr, NCDE_ERR_INVALID, "solve_bwd: gW[n_layers] (gate) is null");
            unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dW3acc, db3acc, gW[m.n_layers], gbias[m.n_layers],
                                                                                 pl.H, pl.C, pl.Cp, pl.Hg, pl.Npad, pl.DF,
                                                                                 pl.DFP, pl.Np, pl.n_bt, 2, 0);
            unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dW3acc, db3acc, gW[pl.F], gbias[pl.F], pl.H,
                                                                                 pl.C, pl.Cp, pl.Hg, pl.Npad, pl.DF, pl.DFP,
                                                                                 pl.Np, pl.n_bt, 2, 1);
            launches += 2;
        } else {
            unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dW3acc, db3acc, gW[pl.F], gbias[pl.F], pl.H,
                                                                                 pl.C, pl.Cp, pl.Hg, pl.Npad, pl.DF, pl.DFP,
                                                                                 pl.Np, pl.n_bt);
            ++launches;
        }
    }
    NCDE_CUDA_OK(cudaGetLastError());
    if (launches_out) *launches_out = launches;
    return NCDE_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// dopri5 forward.  Step control runs on the device (adaptive_kernels.cuh); the host enqueues attempts in chunks and,
// between chunks, looks at a completion flag that the GPU wrote one chunk earlier (so the GPU never waits for the
// host).  stats (device int64[4], nullable) receives attempted, accepted, nfe, flags.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ncde_solve_adaptive_fwd(const ncde_problem_t* p, const float* z0, float* z_out, void* workspace,
                                       size_t workspace_bytes, int64_t* stats, int64_t* launches_out, void* stream) {
    ncde::DeviceGuard device_guard(z0);
    NCDE_REQUIRE(p && z0 && z_out && workspace, NCDE_ERR_INVALID, "solve_adaptive_fwd: null pointer");
    NCDE_REQUIRE(p->method == NCDE_DOPRI5, NCDE_ERR_INVALID, "solve_adaptive_fwd: method must be dopri5");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc != NCDE_OK) return rc;
    const ncde_adaptive_t& ad = p->adaptive;
    NCDE_REQUIRE(ad.n_out >= 1 && ad.out_t, NCDE_ERR_INVALID, "solve_adaptive_fwd: output times missing");
    for (int64_t j = 1; j < ad.n_out; ++j)
        NCDE_REQUIRE(ad.out_t[j] > ad.out_t[j - 1], NCDE_ERR_INVALID, "t must be strictly increasing");
    NCDE_REQUIRE(ad.max_attempts >= 1, NCDE_ERR_INVALID, "solve_adaptive_fwd: max_attempts must be positive");
    NCDE_REQUIRE(p->path.K >= 2 && p->path.knots && p->path.coeffs, NCDE_ERR_INVALID, "solve: bad path");
    NCDE_REQUIRE(workspace_bytes >= adaptive_workspace_floats(pl, ad.n_out) * 4, NCDE_ERR_WORKSPACE,
                 "solve_adaptive_fwd: workspace of %zu bytes is too small", workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t launches = 0;
    const size_t nHB = (size_t)pl.H * pl.Bp;

    Carver cv{(char*)workspace, 0, workspace_bytes};
    float* wpack = cv.take(pl.wpack_floats);
    float* yT = cv.take(nHB);
    float* y1T = cv.take(nHB);
    float* kT[7];
    for (int i = 0; i < 7; ++i) kT[i] = cv.take(nHB);
    float* stage = cv.take(pl.stage_floats);
    AdaptCtrl* ctrl = (AdaptCtrl*)cv.take(sizeof(AdaptCtrl) / 4 + 1);
    const int nblocks = 128;
    double* partials = (double*)cv.take(2 * nblocks * 2);
    double* d_out_t = (double*)cv.take((size_t)ad.n_out * 2);

    rc = pack_weights(p, pl, wpack, 0, st, &launches);
    if (rc != NCDE_OK) return rc;
    const bool use_tc = pl.tc != 0;
    if (use_tc) rc = opt_in_smem(tc_field_fwd_kernel, pl.fwd_smem);
    else if (pl.TM == 8) rc = opt_in_smem(field_fwd_kernel<8>, pl.fwd_smem);
    else rc = opt_in_smem(field_fwd_kernel<4>, pl.fwd_smem);
    if (rc != NCDE_OK) return rc;

    const dim3 tb(32, 8), tg((unsigned)ceil_div(pl.B, 32), (unsigned)ceil_div(pl.H, 32));
    to_feature_major_kernel<<<tg, tb, 0, st>>>(z0, yT, pl.B, pl.Bp, pl.H);
    ++launches;
    NCDE_CUDA_OK(cudaMemcpyAsync(z_out, z0, (size_t)pl.B * pl.H * 4, cudaMemcpyDeviceToDevice, st));
    NCDE_CUDA_OK(cudaMemcpyAsync(d_out_t, ad.out_t, (size_t)ad.n_out * 8, cudaMemcpyHostToDevice, st));

    AdaptParams ap;
    ap.t0 = ad.out_t[0]; ap.rtol = ad.rtol; ap.atol = ad.atol; ap.min_step = ad.min_step; ap.max_step = ad.max_step;
    ap.first_step = ad.first_step; ap.safety = ad.safety; ap.ifactor = ad.ifactor; ap.dfactor = ad.dfactor;
    ap.max_attempts = ad.max_attempts; ap.n_out = (int)ad.n_out; ap.time_sign = 1; ap.keep_counters = 0;
    adapt_init_kernel<<<1, 32, 0, st>>>(ctrl, ap);
    ++launches;

    HiddenFwdArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.B = pl.B; ha.Bp = pl.Bp; ha.H = pl.H; ha.C = pl.C; ha.Cp = pl.Cp; ha.R = pl.R; ha.F = pl.F; ha.Dmax = pl.Dmax;
    for (int l = 0; l <= pl.F; ++l) ha.D[l] = pl.D[l];
    for (int l = 0; l < pl.F; ++l) {
        ha.ldw[l] = pl.ldw[l]; ha.act[l] = p->mlp.act[l];
        ha.WT[l] = wpack + pl.off_WT[l]; ha.bp[l] = wpack + pl.off_bp[l];
        ha.wsm_off[l] = (int)(pl.off_WT[l] - pl.off_WT[0]);
    }
    ha.w_in_smem = pl.w_in_smem; ha.wsm_floats = (int)round_up(pl.wt_floats, 4);
    rc = opt_in_smem(hidden_fwd_kernel, pl.hid_smem_fwd);
    if (rc != NCDE_OK) return rc;
    ha.path.kind = p->path.kind; ha.path.K = (int)p->path.K; ha.path.knots = p->path.knots;
    ha.path.coeffs = p->path.coeffs; ha.path.derivs = p->path.derivs;
    ha.path.match = p->path.match; ha.path.match_terms = p->path.match_terms; ha.path.match_eps = p->path.match_eps;
    for (int i = 0; i < 7; ++i) ha.kT[i] = kT[i];
    ha.combine = COMBINE_LINEAR;
    ha.yT = yT;
    ha.ctrl = ctrl;
    ha.KP = pl.KP;
    for (int l = 0; l <= pl.F; ++l) ha.actT[l] = stage + pl.act_off[l];
    ha.dXT = stage + pl.dx_off;
    ha.abf = use_tc ? (__nv_bfloat16*)(stage + pl.abf_off) : nullptr;

    FieldArgs fa;
    fill_field_args(fa, pl, wpack);
    fa.actT = stage + pl.act_off[pl.F]; fa.dXT = stage + pl.dx_off; fa.ctrl = ctrl;
    TcFieldArgs ta;
    fill_tc_args(ta, pl, wpack);
    ta.abf = (const __nv_bfloat16*)(stage + pl.abf_off); ta.dXT = stage + pl.dx_off; ta.ctrl = ctrl;
    ha.dx_row_major = use_tc ? 1 : 0;
    TcMapSet ms;
    if (use_tc) {
        rc = build_tc_maps(pl, wpack, &ms, stage + pl.abf_off, 0, 1, stage + pl.dx_off, 0, 1);
        if (rc != NCDE_OK) return rc;
    }

    // one vector-field evaluation: stage input from tab[tab_index], result into kT[k_out]
    cudaStream_t cur = st;   // the stream the evaluation kernels go to (the capture stream while the attempt body is recorded)
    auto eval = [&](int tab_index, int k_out, float* stage_input_T) -> int {
        ha.tab_index = tab_index;
        ha.actT[0] = stage_input_T ? stage_input_T : stage + pl.act_off[0];
        NCDE_CUDA_OK(launch_pdl(hidden_fwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_fwd, cur, ha));
        if (pl.F == 0) fa.actT = ha.actT[0];
        if (use_tc) {
            ta.koutT = kT[k_out];
            { const int rc_tc = launch_tc_fwd(pl, ta, ms, cur); if (rc_tc != NCDE_OK) return rc_tc; }
        } else {
            fa.koutT = kT[k_out];
            const dim3 fg(pl.n_hg, pl.n_bt);
            if (pl.TM == 8) NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<8>, fg, dim3(kThreads), pl.fwd_smem, cur, fa));
            else NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<4>, fg, dim3(kThreads), pl.fwd_smem, cur, fa));
        }
        launches += 2;
        return NCDE_OK;
    };

    // f0 = f(t0, y0)
    rc = eval(0, 0, nullptr);
    if (rc != NCDE_OK) return rc;
    const double n_elems = (double)pl.B * pl.H;
    if (!(ad.first_step > 0)) {
        // _select_initial_step (misc.py:32-71)
        NCDE_CUDA_OK(launch_pdl(adapt_norm_kernel, dim3(nblocks), dim3(256), 0, st, (const AdaptCtrl*)ctrl, 0, (const float*)yT,
                                (const float*)kT[0], (const float*)kT[0], pl.B, pl.Bp, pl.H, partials));
        NCDE_CUDA_OK(launch_pdl(adapt_init_step1_kernel, dim3(1), dim3(32), 0, st, ctrl, (const double*)partials, nblocks, n_elems));
        launches += 2;
        rc = eval(NCDE_MAX_STAGES, 1, nullptr);
        if (rc != NCDE_OK) return rc;
        NCDE_CUDA_OK(launch_pdl(adapt_norm_kernel, dim3(nblocks), dim3(256), 0, st, (const AdaptCtrl*)ctrl, 1, (const float*)yT,
                                (const float*)kT[0], (const float*)kT[1], pl.B, pl.Bp, pl.H, partials));
        NCDE_CUDA_OK(launch_pdl(adapt_init_step2_kernel, dim3(1), dim3(32), 0, st, ctrl, (const double*)partials, nblocks, n_elems));
        launches += 2;
    }

    DopriArgs da;
    memset(&da, 0, sizeof(da));
    da.ctrl = ctrl; da.B = pl.B; da.Bp = pl.Bp; da.H = pl.H; da.yT = yT; da.y1T = y1T;
    for (int i = 0; i < 7; ++i) da.kT[i] = kT[i];
    da.out_t = d_out_t; da.z_out = z_out; da.partials = partials; da.nblocks = nblocks;

    auto attempt = [&](cudaStream_t q) -> int {
        cur = q;
        for (int i = 1; i <= 6; ++i) {
            const int rc_e = eval(i, i, i == 6 ? y1T : nullptr);
            if (rc_e != NCDE_OK) { cur = st; return rc_e; }
        }
        cur = st;
        NCDE_CUDA_OK(launch_pdl(dopri_err_kernel, dim3(nblocks), dim3(256), 0, q, da));
        NCDE_CUDA_OK(launch_pdl(dopri_ctrl_kernel, dim3(1), dim3(32), 0, q, da));
        NCDE_CUDA_OK(launch_pdl(dopri_accept_kernel, tg, tb, 0, q, da));
        launches += 3;
        return NCDE_OK;
    };
    if (!dopri_host_poll()) {
        // the attempt loop runs on the device: one graph launch, no host synchronisation (graph_while_not_done)
        rc = graph_while_not_done(st, ctrl, attempt);
        if (rc != NCDE_OK) return rc;
    } else {
        // completion polling: pinned flag + event per chunk, always one chunk of look-ahead
        static int* h_done = nullptr;
        static cudaEvent_t ev[2] = {nullptr, nullptr};
        if (!h_done) {
            NCDE_CUDA_OK(cudaHostAlloc((void**)&h_done, 2 * sizeof(int), cudaHostAllocDefault));
            NCDE_CUDA_OK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
            NCDE_CUDA_OK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        }
        h_done[0] = h_done[1] = 0;
        const int64_t chunk = 32;
        int64_t enq = 0;
        int n_chunks = 0;
        while (enq < ad.max_attempts) {
            const int64_t upto = enq + chunk < ad.max_attempts ? enq + chunk : ad.max_attempts;
            for (; enq < upto; ++enq) {
                for (int i = 1; i <= 6; ++i) {
                    rc = eval(i, i, i == 6 ? y1T : nullptr);
                    if (rc != NCDE_OK) return rc;
                }
                NCDE_CUDA_OK(launch_pdl(dopri_err_kernel, dim3(nblocks), dim3(256), 0, st, da));
                NCDE_CUDA_OK(launch_pdl(dopri_ctrl_kernel, dim3(1), dim3(32), 0, st, da));
                NCDE_CUDA_OK(launch_pdl(dopri_accept_kernel, tg, tb, 0, st, da));
                launches += 3;
            }
            const int slot = n_chunks & 1;
            if (n_chunks >= 1) {
                // wait for the flag written at the end of the PREVIOUS chunk (the chunk just enqueued keeps the GPU busy)
                NCDE_CUDA_OK(cudaEventSynchronize(ev[slot ^ 1]));
                if (h_done[slot ^ 1]) break;
            }
            NCDE_CUDA_OK(cudaMemcpyAsync(&h_done[slot], &ctrl->done, sizeof(int), cudaMemcpyDeviceToHost, st));
            NCDE_CUDA_OK(cudaEventRecord(ev[slot], st));
            ++n_chunks;
        }
    }
    if (stats) {
        // attempted, accepted, nfe are consecutive int64 fields of the control block; flags follows max_attempts
        NCDE_CUDA_OK(cudaMemcpyAsync(stats, &ctrl->attempted, 3 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        NCDE_CUDA_OK(cudaMemsetAsync(stats + 3, 0, sizeof(int64_t), st));
        NCDE_CUDA_OK(cudaMemcpyAsync(stats + 3, &ctrl->flags, sizeof(int), cudaMemcpyDeviceToDevice, st));
        // stats[4] = initial step size (double bits), stats[5..6] = h0, d0, d1, d2 (float bits) of its selection
        NCDE_CUDA_OK(cudaMemcpyAsync(stats + 4, &ctrl->dt_init, sizeof(double), cudaMemcpyDeviceToDevice, st));
        NCDE_CUDA_OK(cudaMemcpyAsync(stats + 5, &ctrl->h0, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        NCDE_CUDA_OK(cudaMemcpyAsync(stats + 8, &ctrl->trace[0][0], 64 * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    NCDE_CUDA_OK(cudaGetLastError());
    if (launches_out) *launches_out = launches;
    return NCDE_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Continuous adjoint with a fixed-grid method (torchdiffeq/_impl/adjoint.py:36-145 with adjoint_method euler / rk4).
// For every output interval [t_{i-1}, t_i], last to first, the augmented state (y, a, g_theta) is integrated in
// reversed time with the same scheme as the forward pass; every augmented evaluation is
//     hidden_fwd(y) -> field_fwd (f)  and  field_bwd(gk = a) -> hidden_bwd (a^T df/dy) -> hidden_wgrad (a^T df/dtheta).
// The time-gradient component of the reference's augmented state never influences a fixed-grid solve and is omitted.
// ---------------------------------------------------------------------------------------------------------------
struct ThetaLayout {
    int n_slots;                       // unique hidden parameter sets
    size_t off_W[NCDE_MAX_LAYERS], off_b[NCDE_MAX_LAYERS];   // per layer (shared layers share offsets); off_b = SIZE_MAX if no bias
    size_t total;
};

static void theta_layout(const ncde_problem_t* p, const Plan& pl, float* const* gW, float* const* gbias, ThetaLayout* tl) {
    size_t off = 0;
    const int n = pl.F + 1;
    for (int l = 0; l < n; ++l) {
        int first = l;
        for (int j = 0; j < l; ++j) if (gW[j] == gW[l]) { first = j; break; }
        if (first != l) { tl->off_W[l] = tl->off_W[first]; tl->off_b[l] = tl->off_b[first]; continue; }
        tl->off_W[l] = off; off += round_up((size_t)p->mlp.out_dim[l] * p->mlp.in_dim[l], 64);
        if (gbias[l]) { tl->off_b[l] = off; off += round_up((size_t)p->mlp.out_dim[l], 64); } else tl->off_b[l] = (size_t)-1;
    }
    if (pl.gated) {   // the sigmoid head of a gated final layer: slot F + 1, shaped like the last layer
        tl->off_W[n] = off; off += round_up((size_t)p->mlp.out_dim[pl.F] * p->mlp.in_dim[pl.F], 64);
        if (gbias[n]) { tl->off_b[n] = off; off += round_up((size_t)p->mlp.out_dim[pl.F], 64); } else tl->off_b[n] = (size_t)-1;
    }
    tl->total = off;
}

static size_t adjoint_workspace_floats(const ncde_problem_t* p, const Plan& pl) {
    size_t per = 256 / 4;
    const size_t nHB = (size_t)pl.H * pl.Bp;
    size_t n = pl.wpack_floats + per;
    n += (size_t)(3 + 2 * 4) * (nHB + per);            // y, a, a_stage, kf[4], ka[4]
    n += pl.stage_floats + per;
    n += (size_t)pl.n_hg * pl.Bp * pl.DFP + per;
    for (int l = 0; l < pl.F; ++l) n += (size_t)pl.Dp4[l + 1] * pl.Bp + per;
    n += (size_t)pl.n_bt * pl.Np * pl.DFP + per + (size_t)pl.n_bt * pl.Np + per;
    n += (size_t)pl.wg_split * (pl.wr_floats + 64 * pl.F) + (size_t)pl.wg_split * pl.F * 1024 + 2 * per;
    size_t ntheta = 0;
    for (int l = 0; l <= pl.F; ++l) ntheta += round_up((size_t)p->mlp.out_dim[l] * p->mlp.in_dim[l], 64) + round_up(p->mlp.out_dim[l], 64);
    if (pl.gated) ntheta += round_up((size_t)p->mlp.out_dim[pl.F] * p->mlp.in_dim[pl.F], 64) + round_up(p->mlp.out_dim[pl.F], 64);
    if (pl.vf) n += 4 * (size_t)pl.Bp + per;
    n += 5 * (ntheta + per);
    return n;
}

extern "C" size_t ncde_solve_adjoint_workspace_bytes(const ncde_problem_t* p) {
    Plan pl;
    if (!p || make_plan(p, &pl, false, true) != NCDE_OK) return 0;
    return adjoint_workspace_floats(p, pl) * 4 + 4096;
}

extern "C" int ncde_solve_adjoint_bwd(const ncde_problem_t* p, const int64_t* interval_steps, int64_t n_out,
                                      const float* y_out, const float* grad_out, float* grad_z0, float* const* gW,
                                      float* const* gbias, void* workspace, size_t workspace_bytes, int64_t* launches_out,
                                      void* stream) {
    ncde::DeviceGuard device_guard(workspace);
    NCDE_REQUIRE(p && interval_steps && y_out && grad_out && grad_z0 && gW && gbias && workspace, NCDE_ERR_INVALID,
                 "solve_adjoint_bwd: null pointer");
    NCDE_REQUIRE(p->method == NCDE_EULER || p->method == NCDE_RK4_38, NCDE_ERR_UNSUPPORTED,
                 "solve_adjoint_bwd: only fixed-grid adjoint methods (euler, rk4) are implemented");
    NCDE_REQUIRE(n_out >= 1, NCDE_ERR_INVALID, "solve_adjoint_bwd: no outputs");
    Plan pl;
    int rc = make_plan(p, &pl, false, true);
    if (rc != NCDE_OK) return rc;
    NCDE_REQUIRE(workspace_bytes >= adjoint_workspace_floats(p, pl) * 4, NCDE_ERR_WORKSPACE,
                 "solve_adjoint_bwd: workspace of %zu bytes is too small", workspace_bytes);
    NCDE_REQUIRE(p->path.K >= 2 && p->path.knots && p->path.coeffs, NCDE_ERR_INVALID, "solve: bad path");
    const ncde_fixed_grid_t& g = p->grid;
    int64_t total_steps = 0;
    for (int64_t i = 0; i + 1 < n_out; ++i) { NCDE_REQUIRE(interval_steps[i] >= 0, NCDE_ERR_INVALID, "bad interval"); total_steps += interval_steps[i]; }
    NCDE_REQUIRE(total_steps == g.n_steps && (g.n_steps == 0 || (g.stage_t && g.dt)), NCDE_ERR_INVALID,
                 "solve_adjoint_bwd: schedule does not match the intervals");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t launches = 0;
    const ncde_mlp_t& m = p->mlp;
    const int NS = pl.n_stages;
    const size_t nHB = (size_t)pl.H * pl.Bp;
    for (int l = 0; l <= pl.F + (pl.gated ? 1 : 0); ++l)
        NCDE_REQUIRE(gW[l] != nullptr, NCDE_ERR_INVALID, "solve_adjoint_bwd: gW[%d] is null", l);

    Carver cv{(char*)workspace, 0, workspace_bytes};
    float* wpack = cv.take(pl.wpack_floats);
    float* e0 = pl.vf ? cv.take(4 * (size_t)pl.Bp) : nullptr;
    if (e0) {
        fill_e0_kernel<<<(unsigned)ceil_div(4 * (int64_t)pl.Bp, 256), 256, 0, (cudaStream_t)stream>>>(e0, pl.Bp);
        ++launches;
    }
    float* yT = cv.take(nHB);
    float* aT = cv.take(nHB);
    float* a_stage = cv.take(nHB);
    float* kf[4]; float* ka[4];
    for (int i = 0; i < 4; ++i) { kf[i] = cv.take(nHB); ka[i] = cv.take(nHB); }
    float* stage = cv.take(pl.stage_floats);
    float* P = cv.take((size_t)pl.n_hg * pl.Bp * pl.DFP);
    float* dpreT[NCDE_MAX_LAYERS] = {};
    for (int l = 0; l < pl.F; ++l) dpreT[l] = cv.take((size_t)pl.Dp4[l + 1] * pl.Bp);
    const size_t nW3 = (size_t)pl.n_bt * pl.Np * pl.DFP, nb3 = (size_t)pl.n_bt * pl.Np;
    float* dW3acc = cv.take(nW3);
    float* db3acc = cv.take(nb3);
    ThetaLayout tl;
    theta_layout(p, pl, gW, gbias, &tl);
    float* theta = cv.take(tl.total);
    float* ktheta[4];
    for (int i = 0; i < 4; ++i) ktheta[i] = cv.take(tl.total);

    rc = pack_weights(p, pl, wpack, 1, st, &launches);
    if (rc != NCDE_OK) return rc;
    const bool use_tc = pl.tc != 0;
    if (use_tc) { rc = opt_in_smem(tc_field_fwd_kernel, pl.fwd_smem); if (rc == NCDE_OK) rc = (pl.bwd_ew == 16 ? opt_in_smem(tc_field_bwd_kernel<16>, pl.bwd_smem) : opt_in_smem(tc_field_bwd_kernel<8>, pl.bwd_smem)); }
    else { rc = opt_in_field(pl, false); if (rc == NCDE_OK) rc = opt_in_field(pl, true); }
    if (rc == NCDE_OK) rc = opt_in_smem(hidden_fwd_kernel, pl.hid_smem_fwd);
    if (rc == NCDE_OK) rc = opt_in_smem(hidden_bwd_kernel, pl.hid_smem_bwd);
    if (rc != NCDE_OK) return rc;

    // ---- argument blocks ----
    HiddenFwdArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.B = pl.B; ha.Bp = pl.Bp; ha.H = pl.H; ha.C = pl.C; ha.Cp = pl.Cp; ha.R = pl.R; ha.F = pl.F; ha.Dmax = pl.Dmax;
    for (int l = 0; l <= pl.F; ++l) ha.D[l] = pl.D[l];
    for (int l = 0; l < pl.F; ++l) {
        ha.ldw[l] = pl.ldw[l]; ha.act[l] = m.act[l];
        ha.WT[l] = wpack + pl.off_WT[l]; ha.bp[l] = wpack + pl.off_bp[l];
        ha.wsm_off[l] = (int)(pl.off_WT[l] - pl.off_WT[0]);
    }
    ha.w_in_smem = pl.w_in_smem; ha.wsm_floats = (int)round_up(pl.wt_floats, 4);
    ha.path.kind = p->path.kind; ha.path.K = (int)p->path.K; ha.path.knots = p->path.knots;
    ha.path.coeffs = p->path.coeffs; ha.path.derivs = p->path.derivs;
    ha.path.match = p->path.match; ha.path.match_terms = p->path.match_terms; ha.path.match_eps = p->path.match_eps;
    for (int i = 0; i < NS; ++i) ha.kT[i] = kf[i];
    ha.yT = yT; ha.KP = pl.KP; ha.comb_sign = -1.f;
    for (int l = 0; l <= pl.F; ++l) ha.actT[l] = stage + pl.act_off[l];
    ha.dXT = stage + pl.dx_off;
    if (pl.vf) {   // X(t) / dX/dt(t) is evaluated by hidden_fwd into the first layer's input; the contraction runs against dX/dt = 1
        ha.dXT = nullptr; ha.vf = pl.vf; ha.n_u = pl.PC; ha.uT = nullptr;
    }
    ha.abf = use_tc ? (__nv_bfloat16*)(stage + pl.abf_off) : nullptr;

    FieldArgs fa;
    fill_field_args(fa, pl, wpack);
    fa.actT = stage + pl.act_off[pl.F]; fa.dXT = pl.vf ? e0 : stage + pl.dx_off;
    fa.P = P; fa.dW3acc = dW3acc; fa.db3acc = db3acc; fa.gkT = a_stage;
    TcFieldArgs ta;
    fill_tc_args(ta, pl, wpack);
    ta.abf = (const __nv_bfloat16*)(stage + pl.abf_off); ta.dXT = stage + pl.dx_off;
    ta.P = P; ta.dW3acc = dW3acc; ta.db3acc = db3acc; ta.gkT = a_stage;
    ha.dx_row_major = use_tc ? 1 : 0;
    TcMapSet ms;
    if (use_tc) {
        rc = build_tc_maps(pl, wpack, &ms, stage + pl.abf_off, 0, 1, stage + pl.dx_off, 0, 1);
        if (rc != NCDE_OK) return rc;
    }

    HiddenBwdArgs hb;
    memset(&hb, 0, sizeof(hb));
    hb.B = pl.B; hb.Bp = pl.Bp; hb.H = pl.H; hb.R = pl.R; hb.F = pl.F; hb.Dmax = pl.Dmax; hb.DFP = pl.DFP; hb.n_hg = pl.n_hg;
    for (int l = 0; l <= pl.F; ++l) { hb.D[l] = pl.D[l]; hb.actT[l] = stage + pl.act_off[l]; }
    for (int l = 0; l < pl.F; ++l) {
        hb.act[l] = m.act[l]; hb.W[l] = wpack + pl.off_WR[l]; hb.ldi[l] = pl.ldi[l]; hb.dpreT[l] = dpreT[l];
        hb.wsm_off[l] = (int)(pl.off_WR[l] - pl.off_WR[0]);
    }
    hb.w_in_smem = pl.w_in_smem; hb.wsm_floats = (int)round_up(pl.wr_floats, 4);
    hb.P = P;

    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.B = pl.B; wa.Bp = pl.Bp; wa.n_split = pl.wg_split; wa.rows_per_split = pl.wg_rows; wa.n_stage = 1;
    int total_tiles = 0;
    int slot_layer[NCDE_MAX_LAYERS];
    for (int l = 0; l < pl.F; ++l) {
        int sidx = -1;
        for (int s2 = 0; s2 < wa.n_slots; ++s2) if (slot_layer[s2] >= 0 && gW[slot_layer[s2]] == gW[l]) sidx = s2;
        if (sidx < 0) {
            sidx = wa.n_slots++;
            slot_layer[sidx] = l;
            wa.Dout[sidx] = m.out_dim[l]; wa.Din[sidx] = m.in_dim[l];
        }
        wa.lay[sidx][wa.n_lay[sidx]++] = l;
        wa.dpreT[0][l] = dpreT[l];
        wa.actT[0][l] = stage + pl.act_off[l];
    }
    size_t gwp_floats[NCDE_MAX_LAYERS], gbp_floats[NCDE_MAX_LAYERS];
    for (int s2 = 0; s2 < wa.n_slots; ++s2) {
        wa.tile_begin[s2] = total_tiles;
        total_tiles += (int)(ceil_div(wa.Dout[s2], kWgTile) * ceil_div(wa.Din[s2], kWgTile));
        gwp_floats[s2] = (size_t)pl.wg_split * wa.Dout[s2] * wa.Din[s2];
        gbp_floats[s2] = (size_t)pl.wg_split * wa.Dout[s2];
        wa.gWp[s2] = cv.take(gwp_floats[s2]);
        wa.gbp[s2] = cv.take(gbp_floats[s2]);
    }
    wa.tile_begin[wa.n_slots] = total_tiles;
    NCDE_REQUIRE(cv.used <= workspace_bytes, NCDE_ERR_WORKSPACE, "solve_adjoint_bwd: workspace accounting error");

    const dim3 tb(32, 8), tg((unsigned)ceil_div(pl.B, 32), (unsigned)ceil_div(pl.H, 32));
    const unsigned ew_grid = (unsigned)ceil_div((int64_t)nHB, 256);
    const unsigned th_grid = (unsigned)ceil_div((int64_t)tl.total, 256);
    static const int rk4_combine[4] = {COMBINE_Y, COMBINE_RK4_S2, COMBINE_RK4_S3, COMBINE_RK4_S4};

    // initial augmented state: y = y(t_T), a = dL/dy(t_T), g_theta = 0
    to_feature_major_kernel<<<tg, tb, 0, st>>>(y_out + (size_t)(n_out - 1) * pl.B * pl.H, yT, pl.B, pl.Bp, pl.H);
    to_feature_major_kernel<<<tg, tb, 0, st>>>(grad_out + (size_t)(n_out - 1) * pl.B * pl.H, aT, pl.B, pl.Bp, pl.H);
    launches += 2;
    NCDE_CUDA_OK(cudaMemsetAsync(theta, 0, tl.total * 4, st));

    int64_t step_idx = 0;
    for (int64_t iv = 0; iv + 1 < n_out; ++iv) {
        const int64_t i_hi = n_out - 1 - iv;  // interval [t_{i_hi-1}, t_{i_hi}]
        for (int64_t s = 0; s < interval_steps[iv]; ++s, ++step_idx) {
            const float dt = g.dt[step_idx];
            for (int sg = 0; sg < NS; ++sg) {
                const int mode = p->method == NCDE_RK4_38 ? rk4_combine[sg] : COMBINE_Y;
                // y stage input (reversed time: k_y = -f) -> activations, dX/dt at the stage time, f
                ha.combine = mode; ha.dt = dt; ha.path.t = g.stage_t[step_idx * NS + sg];
                NCDE_CUDA_OK(launch_pdl(hidden_fwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_fwd, st, ha));
                if (use_tc) {
                    ta.koutT = kf[sg];
                    { const int rc_tc = launch_tc_fwd(pl, ta, ms, st); if (rc_tc != NCDE_OK) return rc_tc; }
                } else {
                    fa.koutT = kf[sg];
                    NCDE_CUDA_OK(launch_field(pl, fa, false, st));
                }
                // adjoint stage input a_s = a + increment(ka)
                AugCombineArgs ca;
                memset(&ca, 0, sizeof(ca));
                ca.n = (int64_t)nHB; ca.combine = mode; ca.dt = dt; ca.sign = 1.f; ca.base = aT; ca.out = a_stage;
                for (int j = 0; j < 4; ++j) ca.k[j] = ka[j];
                NCDE_CUDA_OK(launch_pdl(aug_combine_kernel, dim3(ew_grid), dim3(256), 0, st, ca));
                launches += 3;
                // vector-Jacobian products with a_s
                NCDE_CUDA_OK(cudaMemsetAsync(dW3acc, 0, nW3 * 4, st));
                NCDE_CUDA_OK(cudaMemsetAsync(db3acc, 0, nb3 * 4, st));
                NCDE_CUDA_OK(cudaMemsetAsync(ktheta[sg], 0, tl.total * 4, st));
                for (int s2 = 0; s2 < wa.n_slots; ++s2) {
                    NCDE_CUDA_OK(cudaMemsetAsync(wa.gWp[s2], 0, gwp_floats[s2] * 4, st));
                    NCDE_CUDA_OK(cudaMemsetAsync(wa.gbp[s2], 0, gbp_floats[s2] * 4, st));
                }
                if (use_tc) { const int rc_tc = launch_tc_bwd(pl, ta, ms, st); if (rc_tc != NCDE_OK) return rc_tc; }
                else NCDE_CUDA_OK(launch_field(pl, fa, true, st));
                hb.dz_out = ka[sg];
                NCDE_CUDA_OK(launch_pdl(hidden_bwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_bwd, st, hb));
                launches += 2;
                if (pl.F > 0) {
                    for (int s2 = 0; s2 < wa.n_slots; ++s2) {
                        const int l0 = slot_layer[s2];
                        wa.gW[s2] = ktheta[sg] + tl.off_W[l0];
                        wa.gb[s2] = tl.off_b[l0] == (size_t)-1 ? nullptr : ktheta[sg] + tl.off_b[l0];
                    }
                    NCDE_CUDA_OK(launch_pdl(hidden_wgrad_kernel, dim3(total_tiles, pl.wg_split), dim3(kThreads), 0, st, wa));
                    int nmax = 0;
                    for (int s2 = 0; s2 < wa.n_slots; ++s2) nmax = wa.Dout[s2] * wa.Din[s2] > nmax ? wa.Dout[s2] * wa.Din[s2] : nmax;
                    hidden_wgrad_reduce_kernel<<<dim3((unsigned)ceil_div(nmax, 256), wa.n_slots), 256, 0, st>>>(wa);
                    launches += 2;
                }
                {
                    const int64_t n = (int64_t)pl.H * pl.C * pl.DF;
                    if (pl.gated) {   // interleaved columns: 2c = sigmoid head (slot F + 1), 2c + 1 = tanh head (slot F)
                        unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
                            dW3acc, db3acc, ktheta[sg] + tl.off_W[pl.F + 1],
                            tl.off_b[pl.F + 1] == (size_t)-1 ? nullptr : ktheta[sg] + tl.off_b[pl.F + 1], pl.H, pl.C, pl.Cp, pl.Hg,
                            pl.Npad, pl.DF, pl.DFP, pl.Np, pl.n_bt, 2, 0);
                        ++launches;
                    }
                    unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
                        dW3acc, db3acc, ktheta[sg] + tl.off_W[pl.F],
                        tl.off_b[pl.F] == (size_t)-1 ? nullptr : ktheta[sg] + tl.off_b[pl.F], pl.H, pl.C, pl.Cp, pl.Hg, pl.Npad,
                        pl.DF, pl.DFP, pl.Np, pl.n_bt, pl.gated ? 2 : 1, pl.gated ? 1 : 0);
                    ++launches;
                }
            }
            // advance the augmented state over this step
            aug_advance_kernel<<<ew_grid, 256, 0, st>>>(yT, kf[0], kf[1], kf[2], kf[3], p->method, dt, -1.f, (int64_t)nHB);
            aug_advance_kernel<<<ew_grid, 256, 0, st>>>(aT, ka[0], ka[1], ka[2], ka[3], p->method, dt, 1.f, (int64_t)nHB);
            aug_advance_kernel<<<th_grid, 256, 0, st>>>(theta, ktheta[0], ktheta[1], ktheta[2], ktheta[3], p->method, dt, 1.f,
                                                       (int64_t)tl.total);
            launches += 3;
        }
        // adjoint.py:131-133: restart from the stored forward state, add the gradient arriving at t_{i-1}
        to_feature_major_kernel<<<tg, tb, 0, st>>>(y_out + (size_t)(i_hi - 1) * pl.B * pl.H, yT, pl.B, pl.Bp, pl.H);
        add_out_grad_kernel<<<tg, tb, 0, st>>>(aT, grad_out + (size_t)(i_hi - 1) * pl.B * pl.H, 1.f, pl.B, pl.Bp, pl.H);
        launches += 2;
    }
    from_feature_major_kernel<<<tg, tb, 0, st>>>(aT, nullptr, grad_z0, pl.B, pl.Bp, pl.H);
    ++launches;
    for (int l = 0; l <= pl.F + (pl.gated ? 1 : 0); ++l) {
        bool first = true;
        for (int j = 0; j < l; ++j) if (gW[j] == gW[l]) first = false;
        if (!first) continue;
        const int ls = l > pl.F ? pl.F : l;   // the gate (slot F + 1) is shaped like the last layer
        const int64_t nw = (int64_t)m.out_dim[ls] * m.in_dim[ls];
        axpy_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, st>>>(gW[l], theta + tl.off_W[l], nw);
        ++launches;
        if (gbias[l]) {
            axpy_kernel<<<(unsigned)ceil_div((int64_t)m.out_dim[ls], 256), 256, 0, st>>>(gbias[l], theta + tl.off_b[l], m.out_dim[ls]);
            ++launches;
        }
    }
    NCDE_CUDA_OK(cudaGetLastError());
    if (launches_out) *launches_out = launches;
    return NCDE_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Continuous adjoint with dopri5 (adjoint.py:36-145 with adjoint_method = dopri5).  One adaptive solve of the augmented
// state per output interval, last interval first, in reversed time; the controller (adaptive_kernels.cuh) uses the
// reference's mixed norm over (y, a, every parameter-gradient tensor).  p->adaptive holds the ADJOINT tolerances and
// options and the forward output times.  The host polls a completion flag after every chunk of 4 attempts.
// ---------------------------------------------------------------------------------------------------------------
static size_t adjoint_adaptive_workspace_floats(const ncde_problem_t* p, const Plan& pl, int64_t n_out) {
    size_t per = 256 / 4;
    const size_t nHB = (size_t)pl.H * pl.Bp;
    size_t n = adjoint_workspace_floats(p, pl);
    n += (size_t)(2 + 2 * 3) * (nHB + per);    // y1, a_out, extra kf[3], ka[3]
    size_t ntheta = 0;
    for (int l = 0; l <= pl.F; ++l) ntheta += round_up((size_t)p->mlp.out_dim[l] * p->mlp.in_dim[l], 64) + round_up(p->mlp.out_dim[l], 64);
    n += 5 * (ntheta + per);                   // ktheta[4..6], theta1, theta_out
    n += sizeof(AdaptCtrl) / 4 + per + (size_t)(kMaxSeg + 1) * 2 * 128 * 2 + per + (size_t)n_out * 4 + per;
    n += nHB + (size_t)pl.Cp * pl.Bp + 3 * per;   // time-gradient component: q, d2X/dt2, scalars
    return n;
}

extern "C" size_t ncde_solve_adjoint_adaptive_workspace_bytes(const ncde_problem_t* p) {
    Plan pl;
    if (!p || make_plan(p, &pl) != NCDE_OK) return 0;
    return adjoint_adaptive_workspace_floats(p, pl, p->adaptive.n_out) * 4 + 4096;
}

extern "C" int ncde_solve_adjoint_adaptive_bwd(const ncde_problem_t* p, const float* y_out, const float* grad_out,
                                               float* grad_z0, float* const* gW, float* const* gbias, void* workspace,
                                               size_t workspace_bytes, int64_t* stats, int64_t* launches_out, void* stream) {
    ncde::DeviceGuard device_guard(y_out);
    NCDE_REQUIRE(p && y_out && grad_out && grad_z0 && gW && gbias && workspace, NCDE_ERR_INVALID,
                 "solve_adjoint_adaptive_bwd: null pointer");
    NCDE_REQUIRE(p->method == NCDE_DOPRI5, NCDE_ERR_INVALID, "solve_adjoint_adaptive_bwd: method must be dopri5");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc != NCDE_OK) return rc;
    const ncde_adaptive_t& ad = p->adaptive;
    const int64_t n_out = ad.n_out;
    NCDE_REQUIRE(n_out >= 1 && ad.out_t && ad.max_attempts >= 1, NCDE_ERR_INVALID, "solve_adjoint_adaptive_bwd: bad adaptive block");
    NCDE_REQUIRE(workspace_bytes >= adjoint_adaptive_workspace_floats(p, pl, n_out) * 4, NCDE_ERR_WORKSPACE,
                 "solve_adjoint_adaptive_bwd: workspace of %zu bytes is too small", workspace_bytes);
    NCDE_REQUIRE(p->path.K >= 2 && p->path.knots && p->path.coeffs, NCDE_ERR_INVALID, "solve: bad path");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t launches = 0;
    const ncde_mlp_t& m = p->mlp;
    const size_t nHB = (size_t)pl.H * pl.Bp;
    for (int l = 0; l <= pl.F; ++l) NCDE_REQUIRE(gW[l] != nullptr, NCDE_ERR_INVALID, "solve_adjoint_adaptive_bwd: gW[%d] is null", l);

    Carver cv{(char*)workspace, 0, workspace_bytes};
    float* wpack = cv.take(pl.wpack_floats);
    float* yT = cv.take(nHB);
    float* y1T = cv.take(nHB);
    float* aT = cv.take(nHB);
    float* a_stage = cv.take(nHB);
    float* a_out = cv.take(nHB);
    float* kf[7]; float* ka[7];
    for (int i = 0; i < 7; ++i) { kf[i] = cv.take(nHB); ka[i] = cv.take(nHB); }
    float* stage = cv.take(pl.stage_floats);
    float* P = cv.take((size_t)pl.n_hg * pl.Bp * pl.DFP);
    float* dpreT[NCDE_MAX_LAYERS] = {};
    for (int l = 0; l < pl.F; ++l) dpreT[l] = cv.take((size_t)pl.Dp4[l + 1] * pl.Bp);
    const size_t nW3 = (size_t)pl.n_bt * pl.Np * pl.DFP, nb3 = (size_t)pl.n_bt * pl.Np;
    float* dW3acc = cv.take(nW3);
    float* db3acc = cv.take(nb3);
    ThetaLayout tl;
    theta_layout(p, pl, gW, gbias, &tl);
    float* theta = cv.take(tl.total);
    float* theta1 = cv.take(tl.total);
    float* theta_out = cv.take(tl.total);
    float* ktheta[7];
    for (int i = 0; i < 7; ++i) ktheta[i] = cv.take(tl.total);
    AdaptCtrl* ctrl = (AdaptCtrl*)cv.take(sizeof(AdaptCtrl) / 4 + 1);
    const int nblocks = 128;
    double* partials = (double*)cv.take((size_t)(kMaxSeg + 1) * 2 * nblocks * 2);
    double* d_tau = (double*)cv.take((size_t)n_out * 4);
    // time-gradient component vjp_t of the augmented state (adjoint.py:64,80-100).  The reference always integrates it
    // and it enters the error norm as |.|; dX/dt of a linear path does not depend on t, so it stays exactly zero there.
    const bool has_vt = p->path.kind != NCDE_PATH_LINEAR;
    float* qT = cv.take(nHB);
    float* ddXT = cv.take((size_t)pl.Cp * pl.Bp);
    float* vts = cv.take(64);   // [0] vt, [1] candidate, [2] dense output at the interval end, [8..14] stage derivatives
    float* kvt[7];
    for (int i = 0; i < 7; ++i) kvt[i] = vts + 8 + i;

    rc = pack_weights(p, pl, wpack, 1, st, &launches);
    if (rc != NCDE_OK) return rc;
    const bool use_tc = pl.tc != 0;
    if (use_tc) { rc = opt_in_smem(tc_field_fwd_kernel, pl.fwd_smem); if (rc == NCDE_OK) rc = (pl.bwd_ew == 16 ? opt_in_smem(tc_field_bwd_kernel<16>, pl.bwd_smem) : opt_in_smem(tc_field_bwd_kernel<8>, pl.bwd_smem)); }
    else if (pl.TM == 8) { rc = opt_in_smem(field_fwd_kernel<8>, pl.fwd_smem); if (rc == NCDE_OK) rc = opt_in_smem(field_bwd_kernel<8>, pl.bwd_smem); }
    else { rc = opt_in_smem(field_fwd_kernel<4>, pl.fwd_smem); if (rc == NCDE_OK) rc = opt_in_smem(field_bwd_kernel<4>, pl.bwd_smem); }
    if (rc == NCDE_OK) rc = opt_in_smem(hidden_fwd_kernel, pl.hid_smem_fwd);
    if (rc == NCDE_OK) rc = opt_in_smem(hidden_bwd_kernel, pl.hid_smem_bwd);
    if (rc != NCDE_OK) return rc;

    HiddenFwdArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.B = pl.B; ha.Bp = pl.Bp; ha.H = pl.H; ha.C = pl.C; ha.Cp = pl.Cp; ha.R = pl.R; ha.F = pl.F; ha.Dmax = pl.Dmax;
    for (int l = 0; l <= pl.F; ++l) ha.D[l] = pl.D[l];
    for (int l = 0; l < pl.F; ++l) {
        ha.ldw[l] = pl.ldw[l]; ha.act[l] = m.act[l];
        ha.WT[l] = wpack + pl.off_WT[l]; ha.bp[l] = wpack + pl.off_bp[l];
        ha.wsm_off[l] = (int)(pl.off_WT[l] - pl.off_WT[0]);
    }
    ha.w_in_smem = pl.w_in_smem; ha.wsm_floats = (int)round_up(pl.wt_floats, 4);
    ha.path.kind = p->path.kind; ha.path.K = (int)p->path.K; ha.path.knots = p->path.knots;
    ha.path.coeffs = p->path.coeffs; ha.path.derivs = p->path.derivs;
    ha.path.match = p->path.match; ha.path.match_terms = p->path.match_terms; ha.path.match_eps = p->path.match_eps;
    for (int i = 0; i < 7; ++i) ha.kT[i] = kf[i];
    ha.yT = yT; ha.KP = pl.KP; ha.comb_sign = -1.f; ha.combine = COMBINE_LINEAR; ha.ctrl = ctrl;
    for (int l = 0; l <= pl.F; ++l) ha.actT[l] = stage + pl.act_off[l];
    ha.dXT = stage + pl.dx_off;
    ha.ddXT = has_vt ? ddXT : nullptr;
    ha.abf = use_tc ? (__nv_bfloat16*)(stage + pl.abf_off) : nullptr;

    FieldArgs fa;
    fill_field_args(fa, pl, wpack);
    fa.actT = stage + pl.act_off[pl.F]; fa.dXT = stage + pl.dx_off;
    fa.P = P; fa.dW3acc = dW3acc; fa.db3acc = db3acc; fa.gkT = a_stage; fa.ctrl = ctrl;
    TcFieldArgs ta;
    fill_tc_args(ta, pl, wpack);
    ta.abf = (const __nv_bfloat16*)(stage + pl.abf_off); ta.dXT = stage + pl.dx_off;
    ta.P = P; ta.dW3acc = dW3acc; ta.db3acc = db3acc; ta.gkT = a_stage; ta.ctrl = ctrl;
    ha.dx_row_major = use_tc ? 1 : 0;
    TcMapSet ms, ms_q;   // ms_q: the same activations with d2X/dt2 in place of dX/dt (time-gradient component)
    if (use_tc) {
        rc = build_tc_maps(pl, wpack, &ms, stage + pl.abf_off, 0, 1, stage + pl.dx_off, 0, 1);
        if (rc == NCDE_OK) rc = build_tc_maps(pl, wpack, &ms_q, stage + pl.abf_off, 0, 1, ddXT, 0, 1);
        if (rc != NCDE_OK) return rc;
    }

    HiddenBwdArgs hb;
    memset(&hb, 0, sizeof(hb));
    hb.B = pl.B; hb.Bp = pl.Bp; hb.H = pl.H; hb.R = pl.R; hb.F = pl.F; hb.Dmax = pl.Dmax; hb.DFP = pl.DFP; hb.n_hg = pl.n_hg;
    for (int l = 0; l <= pl.F; ++l) { hb.D[l] = pl.D[l]; hb.actT[l] = stage + pl.act_off[l]; }
    for (int l = 0; l < pl.F; ++l) {
        hb.act[l] = m.act[l]; hb.W[l] = wpack + pl.off_WR[l]; hb.ldi[l] = pl.ldi[l]; hb.dpreT[l] = dpreT[l];
        hb.wsm_off[l] = (int)(pl.off_WR[l] - pl.off_WR[0]);
    }
    hb.w_in_smem = pl.w_in_smem; hb.wsm_floats = (int)round_up(pl.wr_floats, 4);
    hb.P = P; hb.ctrl = ctrl;

    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.B = pl.B; wa.Bp = pl.Bp; wa.n_split = pl.wg_split; wa.rows_per_split = pl.wg_rows; wa.n_stage = 1; wa.ctrl = ctrl;
    int total_tiles = 0;
    int slot_layer[NCDE_MAX_LAYERS];
    for (int l = 0; l < pl.F; ++l) {
        int sidx = -1;
        for (int s2 = 0; s2 < wa.n_slots; ++s2) if (gW[slot_layer[s2]] == gW[l]) sidx = s2;
        if (sidx < 0) {
            sidx = wa.n_slots++;
            slot_layer[sidx] = l;
            wa.Dout[sidx] = m.out_dim[l]; wa.Din[sidx] = m.in_dim[l];
        }
        wa.lay[sidx][wa.n_lay[sidx]++] = l;
        wa.dpreT[0][l] = dpreT[l];
        wa.actT[0][l] = stage + pl.act_off[l];
    }
    size_t gwp_floats[NCDE_MAX_LAYERS], gbp_floats[NCDE_MAX_LAYERS];
    for (int s2 = 0; s2 < wa.n_slots; ++s2) {
        wa.tile_begin[s2] = total_tiles;
        total_tiles += (int)(ceil_div(wa.Dout[s2], kWgTile) * ceil_div(wa.Din[s2], kWgTile));
        gwp_floats[s2] = (size_t)pl.wg_split * wa.Dout[s2] * wa.Din[s2];
        gbp_floats[s2] = (size_t)pl.wg_split * wa.Dout[s2];
        wa.gWp[s2] = cv.take(gwp_floats[s2]);
        wa.gbp[s2] = cv.take(gbp_floats[s2]);
    }
    wa.tile_begin[wa.n_slots] = total_tiles;
    NCDE_REQUIRE(cv.used <= workspace_bytes, NCDE_ERR_WORKSPACE, "solve_adjoint_adaptive_bwd: workspace accounting error");

    // segments of the mixed norm: y, a, then every parameter-gradient tensor
    struct Seg { size_t off; int64_t n; };
    std::vector<Seg> tsegs;
    for (int l = 0; l <= pl.F; ++l) {
        bool first = true;
        for (int j = 0; j < l; ++j) if (gW[j] == gW[l]) first = false;
        if (!first) continue;
        tsegs.push_back({tl.off_W[l], (int64_t)m.out_dim[l] * m.in_dim[l]});
        if (tl.off_b[l] != (size_t)-1) tsegs.push_back({tl.off_b[l], (int64_t)m.out_dim[l]});
    }
    const int n_seg = 2 + (int)tsegs.size() + (has_vt ? 1 : 0);   // the time-gradient scalar is the last segment
    NCDE_REQUIRE(n_seg <= kMaxSeg, NCDE_ERR_UNSUPPORTED, "too many parameter tensors");
    AugCtrlArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.ctrl = ctrl; ca.partials = partials; ca.nblocks = nblocks; ca.n_seg = n_seg;
    ca.count[0] = ca.count[1] = (double)pl.B * pl.H;
    for (size_t i = 0; i < tsegs.size(); ++i) ca.count[2 + i] = (double)tsegs[i].n;
    if (has_vt) ca.count[n_seg - 1] = 1.0;

    const dim3 tb(32, 8), tg((unsigned)ceil_div(pl.B, 32), (unsigned)ceil_div(pl.H, 32));
    const unsigned ew_grid = (unsigned)ceil_div((int64_t)nHB, 256);
    const unsigned th_grid = (unsigned)ceil_div((int64_t)tl.total, 256);

    // one augmented evaluation: stage inputs from ctrl->tab[tab_index]; results into kf/ka/ktheta[k_out]
    cudaStream_t cur = st;   // the stream the attempt kernels go to (the capture stream while the attempt body is recorded)
    auto aeval = [&](int tab_index, int k_out, float* y_stage_T) -> int {
        ha.tab_index = tab_index;
        ha.actT[0] = y_stage_T ? y_stage_T : stage + pl.act_off[0];
        hb.actT[0] = ha.actT[0];
        wa.actT[0][0] = ha.actT[0];   // the first hidden layer's input is the stage input itself
        if (pl.F == 0) { fa.actT = ha.actT[0]; }
        NCDE_CUDA_OK(launch_pdl(hidden_fwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_fwd, cur, ha));
        const dim3 fg(pl.n_hg, pl.n_bt);
        if (use_tc) {
            ta.koutT = kf[k_out];
            { const int rc_tc = launch_tc_fwd(pl, ta, ms, cur); if (rc_tc != NCDE_OK) return rc_tc; }
        } else {
            fa.koutT = kf[k_out];
            if (pl.TM == 8) NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<8>, fg, dim3(kThreads), pl.fwd_smem, cur, fa));
            else NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<4>, fg, dim3(kThreads), pl.fwd_smem, cur, fa));
        }
        if (has_vt) {
            // q[b,h] = sum_c F(z)[b,h,c] d2X/dt2[b,c]: the same field kernel with the second path derivative
            if (use_tc) {
                ta.koutT = qT; ta.dXT = ddXT;
                { const int rc_tc = launch_tc_fwd(pl, ta, ms_q, cur); if (rc_tc != NCDE_OK) return rc_tc; }
                ta.dXT = stage + pl.dx_off;
            } else {
                fa.koutT = qT; fa.dXT = ddXT;
                if (pl.TM == 8) NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<8>, fg, dim3(kThreads), pl.fwd_smem, cur, fa));
                else NCDE_CUDA_OK(launch_pdl(field_fwd_kernel<4>, fg, dim3(kThreads), pl.fwd_smem, cur, fa));
                fa.dXT = stage + pl.dx_off;
            }
            ++launches;
        }
        AugCombineArgs cb;
        memset(&cb, 0, sizeof(cb));
        cb.n = (int64_t)nHB; cb.combine = COMBINE_LINEAR; cb.sign = 1.f; cb.base = aT; cb.out = a_stage;
        for (int j = 0; j < 7; ++j) cb.k[j] = ka[j];
        cb.ctrl = ctrl; cb.tab_index = tab_index;
        NCDE_CUDA_OK(launch_pdl(aug_combine_kernel, dim3(ew_grid), dim3(256), 0, cur, cb));
        if (has_vt) {
            // d(vjp_t)/dtau = sum_{b,h} a[b,h] q[b,h]
            NCDE_CUDA_OK(launch_pdl(aug_dot_kernel, dim3(1), dim3(1024), 0, cur, (const AdaptCtrl*)ctrl, (const float*)a_stage,
                                    (const float*)qT, pl.B, pl.Bp, pl.H, kvt[k_out]));
            ++launches;
        }
        // zero-fills as kernels (not memset nodes): this sequence is recorded into the body of a conditional graph node
        auto zero = [&](float* ptr, size_t n) {
            zero_f32_kernel<<<(unsigned)ceil_div((int64_t)n, 1024), 256, 0, cur>>>(ptr, (int64_t)n);
        };
        zero(dW3acc, nW3);
        zero(db3acc, nb3);
        zero(ktheta[k_out], tl.total);
        for (int s2 = 0; s2 < wa.n_slots; ++s2) {
            zero(wa.gWp[s2], gwp_floats[s2]);
            zero(wa.gbp[s2], gbp_floats[s2]);
        }
        if (use_tc) { const int rc_tc = launch_tc_bwd(pl, ta, ms, cur); if (rc_tc != NCDE_OK) return rc_tc; }
        else if (pl.TM == 8) NCDE_CUDA_OK(launch_pdl(field_bwd_kernel<8>, fg, dim3(kThreads), pl.bwd_smem, cur, fa));
        else NCDE_CUDA_OK(launch_pdl(field_bwd_kernel<4>, fg, dim3(kThreads), pl.bwd_smem, cur, fa));
        hb.dz_out = ka[k_out];
        NCDE_CUDA_OK(launch_pdl(hidden_bwd_kernel, dim3(pl.n_rt), dim3(kThreads), pl.hid_smem_bwd, cur, hb));
        launches += 5;
        if (pl.F > 0) {
            for (int s2 = 0; s2 < wa.n_slots; ++s2) {
                const int l0 = slot_layer[s2];
                wa.gW[s2] = ktheta[k_out] + tl.off_W[l0];
                wa.gb[s2] = tl.off_b[l0] == (size_t)-1 ? nullptr : ktheta[k_out] + tl.off_b[l0];
            }
            NCDE_CUDA_OK(launch_pdl(hidden_wgrad_kernel, dim3(total_tiles, pl.wg_split), dim3(kThreads), 0, cur, wa));
            int nmax = 0;
            for (int s2 = 0; s2 < wa.n_slots; ++s2) nmax = wa.Dout[s2] * wa.Din[s2] > nmax ? wa.Dout[s2] * wa.Din[s2] : nmax;
            hidden_wgrad_reduce_kernel<<<dim3((unsigned)ceil_div(nmax, 256), wa.n_slots), 256, 0, cur>>>(wa);
            launches += 2;
        }
        const int64_t n = (int64_t)pl.H * pl.C * pl.DF;
        unpack_final_grad_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, cur>>>(
            dW3acc, db3acc, ktheta[k_out] + tl.off_W[pl.F],
            tl.off_b[pl.F] == (size_t)-1 ? nullptr : ktheta[k_out] + tl.off_b[pl.F], pl.H, pl.C, pl.Cp, pl.Hg, pl.Npad, pl.DF,
            pl.DFP, pl.Np, pl.n_bt);
        ++launches;
        return NCDE_OK;
    };
    // mixed-norm partial sums of the initial-step selection: mode 0 -> (s0, f0), mode 1 -> (f0, f1)
    auto norms = [&](int mode, int kb) -> int {
        const int kb0 = 0;
        NCDE_CUDA_OK(launch_pdl(adapt_norm_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, mode, (const float*)yT,
                                (const float*)kf[kb0], (const float*)kf[kb], pl.B, pl.Bp, pl.H, partials + (size_t)0 * 2 * nblocks));
        NCDE_CUDA_OK(launch_pdl(adapt_norm_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, mode, (const float*)aT,
                                (const float*)ka[kb0], (const float*)ka[kb], pl.B, pl.Bp, pl.H, partials + (size_t)1 * 2 * nblocks));
        for (size_t i = 0; i < tsegs.size(); ++i)
            NCDE_CUDA_OK(launch_pdl(adapt_norm_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, mode,
                                    (const float*)(theta + tsegs[i].off), (const float*)(ktheta[kb0] + tsegs[i].off),
                                    (const float*)(ktheta[kb] + tsegs[i].off), (int)tsegs[i].n, (int)tsegs[i].n, 1,
                                    partials + (size_t)(2 + i) * 2 * nblocks));
        if (has_vt)
            NCDE_CUDA_OK(launch_pdl(adapt_norm_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, mode, (const float*)vts,
                                    (const float*)kvt[kb0], (const float*)kvt[kb], 1, 1, 1, partials + (size_t)(n_seg - 1) * 2 * nblocks));
        launches += n_seg;
        return NCDE_OK;
    };

    // reversed output times tau_j = -t_j, per interval [tau_start, tau_target]
    std::vector<double> h_tau((size_t)2 * (n_out > 1 ? n_out - 1 : 1));
    for (int64_t iv = 0; iv + 1 < n_out; ++iv) {
        const int64_t i_hi = n_out - 1 - iv;
        h_tau[2 * iv] = -ad.out_t[i_hi];
        h_tau[2 * iv + 1] = -ad.out_t[i_hi - 1];
        NCDE_REQUIRE(h_tau[2 * iv + 1] > h_tau[2 * iv], NCDE_ERR_INVALID, "t must be strictly increasing");
    }
    if (n_out > 1) NCDE_CUDA_OK(cudaMemcpy(d_tau, h_tau.data(), h_tau.size() * 8, cudaMemcpyHostToDevice));

    to_feature_major_kernel<<<tg, tb, 0, st>>>(y_out + (size_t)(n_out - 1) * pl.B * pl.H, yT, pl.B, pl.Bp, pl.H);
    to_feature_major_kernel<<<tg, tb, 0, st>>>(grad_out + (size_t)(n_out - 1) * pl.B * pl.H, aT, pl.B, pl.Bp, pl.H);
    launches += 2;
    NCDE_CUDA_OK(cudaMemsetAsync(theta, 0, tl.total * 4, st));
    NCDE_CUDA_OK(cudaMemsetAsync(vts, 0, 64 * 4, st));
    NCDE_CUDA_OK(cudaMemsetAsync(ctrl, 0, sizeof(AdaptCtrl), st));

    for (int64_t iv = 0; iv + 1 < n_out; ++iv) {
        const int64_t i_hi = n_out - 1 - iv;
        const double* out_tau = d_tau + 2 * iv;
        ca.out_t = out_tau;
        AdaptParams ap;
        ap.t0 = h_tau[2 * iv]; ap.rtol = ad.rtol; ap.atol = ad.atol; ap.min_step = ad.min_step; ap.max_step = ad.max_step;
        ap.first_step = ad.first_step; ap.safety = ad.safety; ap.ifactor = ad.ifactor; ap.dfactor = ad.dfactor;
        ap.max_attempts = ad.max_attempts; ap.n_out = 2; ap.time_sign = -1; ap.keep_counters = iv > 0 ? 1 : 0;
        adapt_init_kernel<<<1, 32, 0, st>>>(ctrl, ap);
        ++launches;
        // f0 of the augmented system at the interval start
        rc = aeval(0, 0, nullptr);
        if (rc != NCDE_OK) return rc;
        if (!(ad.first_step > 0)) {
            rc = norms(0, 0);
            if (rc != NCDE_OK) return rc;
            NCDE_CUDA_OK(launch_pdl(aug_init_step1_kernel, dim3(1), dim3(32), 0, st, ca));
            rc = aeval(NCDE_MAX_STAGES, 1, nullptr);
            if (rc != NCDE_OK) return rc;
            rc = norms(1, 1);
            if (rc != NCDE_OK) return rc;
            NCDE_CUDA_OK(launch_pdl(aug_init_step2_kernel, dim3(1), dim3(32), 0, st, ca));
            launches += 2;
        }
        auto attempt = [&](cudaStream_t q) -> int {
            cur = q;
            for (int i = 1; i <= 6; ++i) {
                const int rc_e = aeval(i, i, i == 6 ? y1T : nullptr);
                if (rc_e != NCDE_OK) { cur = st; return rc_e; }
            }
            // theta candidate = theta + sum_j beta_6j dt ktheta_j (the 7th stage input of the parameter component)
            AugCombineArgs cb;
            memset(&cb, 0, sizeof(cb));
            cb.n = (int64_t)tl.total; cb.combine = COMBINE_LINEAR; cb.sign = 1.f; cb.base = theta; cb.out = theta1;
            for (int j = 0; j < 7; ++j) cb.k[j] = ktheta[j];
            cb.ctrl = ctrl; cb.tab_index = 6;
            NCDE_CUDA_OK(launch_pdl(aug_combine_kernel, dim3(th_grid), dim3(256), 0, cur, cb));
            if (has_vt) {
                cb.n = 1; cb.base = vts; cb.out = vts + 1;
                for (int j = 0; j < 7; ++j) cb.k[j] = kvt[j];
                NCDE_CUDA_OK(launch_pdl(aug_combine_kernel, dim3(1), dim3(32), 0, cur, cb));
                ++launches;
            }
            // error ratio per segment
            NCDE_CUDA_OK(launch_pdl(aug_err_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, (const float*)yT,
                                    (const float*)y1T, (const float*)kf[0], (const float*)kf[1], (const float*)kf[2],
                                    (const float*)kf[3], (const float*)kf[4], (const float*)kf[5], (const float*)kf[6],
                                    (int64_t)nHB, pl.Bp, pl.B, partials));
            NCDE_CUDA_OK(launch_pdl(aug_err_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, (const float*)aT,
                                    (const float*)a_stage, (const float*)ka[0], (const float*)ka[1], (const float*)ka[2],
                                    (const float*)ka[3], (const float*)ka[4], (const float*)ka[5], (const float*)ka[6],
                                    (int64_t)nHB, pl.Bp, pl.B, partials + (size_t)1 * 2 * nblocks));
            for (size_t i = 0; i < tsegs.size(); ++i) {
                const size_t o = tsegs[i].off;
                NCDE_CUDA_OK(launch_pdl(aug_err_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl,
                                        (const float*)(theta + o), (const float*)(theta1 + o), (const float*)(ktheta[0] + o),
                                        (const float*)(ktheta[1] + o), (const float*)(ktheta[2] + o), (const float*)(ktheta[3] + o),
                                        (const float*)(ktheta[4] + o), (const float*)(ktheta[5] + o), (const float*)(ktheta[6] + o),
                                        tsegs[i].n, 0, 0, partials + (size_t)(2 + i) * 2 * nblocks));
            }
            if (has_vt)
                NCDE_CUDA_OK(launch_pdl(aug_err_kernel, dim3(nblocks), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, (const float*)vts,
                                        (const float*)(vts + 1), (const float*)kvt[0], (const float*)kvt[1], (const float*)kvt[2],
                                        (const float*)kvt[3], (const float*)kvt[4], (const float*)kvt[5], (const float*)kvt[6],
                                        (int64_t)1, 0, 0, partials + (size_t)(n_seg - 1) * 2 * nblocks));
            NCDE_CUDA_OK(launch_pdl(aug_ctrl_kernel, dim3(1), dim3(32), 0, cur, ca));
            NCDE_CUDA_OK(launch_pdl(aug_accept_kernel, dim3(ew_grid), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, yT, (const float*)y1T,
                                    kf[0], (const float*)kf[1], (const float*)kf[2], (const float*)kf[3], (const float*)kf[4],
                                    (const float*)kf[5], (const float*)kf[6], (float*)nullptr, (int64_t)nHB, out_tau));
            NCDE_CUDA_OK(launch_pdl(aug_accept_kernel, dim3(ew_grid), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, aT, (const float*)a_stage,
                                    ka[0], (const float*)ka[1], (const float*)ka[2], (const float*)ka[3], (const float*)ka[4],
                                    (const float*)ka[5], (const float*)ka[6], a_out, (int64_t)nHB, out_tau));
            NCDE_CUDA_OK(launch_pdl(aug_accept_kernel, dim3(th_grid), dim3(256), 0, cur, (const AdaptCtrl*)ctrl, theta, (const float*)theta1,
                                    ktheta[0], (const float*)ktheta[1], (const float*)ktheta[2], (const float*)ktheta[3],
                                    (const float*)ktheta[4], (const float*)ktheta[5], (const float*)ktheta[6], theta_out,
                                    (int64_t)tl.total, out_tau));
            if (has_vt) {
                NCDE_CUDA_OK(launch_pdl(aug_accept_kernel, dim3(1), dim3(32), 0, cur, (const AdaptCtrl*)ctrl, vts, (const float*)(vts + 1),
                                        kvt[0], (const float*)kvt[1], (const float*)kvt[2], (const float*)kvt[3],
                                        (const float*)kvt[4], (const float*)kvt[5], (const float*)kvt[6], vts + 2, (int64_t)1, out_tau));
                ++launches;
            }
            launches += 7 + n_seg;
            cur = st;
            return NCDE_OK;
        };
        if (!dopri_host_poll()) {
            rc = graph_while_not_done(st, ctrl, attempt);     // device-side loop, no host synchronisation
            if (rc != NCDE_OK) return rc;
        } else {
            static int* h_done = nullptr;
            if (!h_done) NCDE_CUDA_OK(cudaHostAlloc((void**)&h_done, sizeof(int), cudaHostAllocDefault));
            int64_t enq = 0;
            bool finished = false;
            while (enq < ad.max_attempts && !finished) {
                for (int c4 = 0; c4 < 4 && enq < ad.max_attempts; ++c4, ++enq) {
                    rc = attempt(st);
                    if (rc != NCDE_OK) return rc;
                }
                NCDE_CUDA_OK(cudaMemcpyAsync(h_done, &ctrl->done, sizeof(int), cudaMemcpyDeviceToHost, st));
                NCDE_CUDA_OK(cudaStreamSynchronize(st));
                finished = *h_done != 0;
            }
        }
        // interval end (adjoint.py:131-133): adjoint and parameter gradients from the dense output at t_{i-1}, the state
        // from the stored forward solution, plus the gradient arriving at t_{i-1}
        NCDE_CUDA_OK(cudaMemcpyAsync(aT, a_out, nHB * 4, cudaMemcpyDeviceToDevice, st));
        NCDE_CUDA_OK(cudaMemcpyAsync(theta, theta_out, tl.total * 4, cudaMemcpyDeviceToDevice, st));
        if (has_vt) NCDE_CUDA_OK(cudaMemcpyAsync(vts, vts + 2, 4, cudaMemcpyDeviceToDevice, st));
        to_feature_major_kernel<<<tg, tb, 0, st>>>(y_out + (size_t)(i_hi - 1) * pl.B * pl.H, yT, pl.B, pl.Bp, pl.H);
        add_out_grad_kernel<<<tg, tb, 0, st>>>(aT, grad_out + (size_t)(i_hi - 1) * pl.B * pl.H, 1.f, pl.B, pl.Bp, pl.H);
        launches += 2;
    }
    from_feature_major_kernel<<<tg, tb, 0, st>>>(aT, nullptr, grad_z0, pl.B, pl.Bp, pl.H);
    ++launches;
    for (int l = 0; l <= pl.F; ++l) {
        bool first = true;
        for (int j = 0; j < l; ++j) if (gW[j] == gW[l]) first = false;
        if (!first) continue;
        const int64_t nw = (int64_t)m.out_dim[l] * m.in_dim[l];
        axpy_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, st>>>(gW[l], theta + tl.off_W[l], nw);
        ++launches;
        if (gbias[l]) {
            axpy_kernel<<<(unsigned)ceil_div((int64_t)m.out_dim[l], 256), 256, 0, st>>>(gbias[l], theta + tl.off_b[l], m.out_dim[l]);
            ++launches;
        }
    }
    if (stats) {
        NCDE_CUDA_OK(cudaMemcpyAsync(stats, &ctrl->attempted, 3 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        NCDE_CUDA_OK(cudaMemsetAsync(stats + 3, 0, sizeof(int64_t), st));
        NCDE_CUDA_OK(cudaMemcpyAsync(stats + 3, &ctrl->flags, sizeof(int), cudaMemcpyDeviceToDevice, st));
        NCDE_CUDA_OK(cudaMemcpyAsync(stats + 8, &ctrl->trace[0][0], 64 * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    NCDE_CUDA_OK(cudaGetLastError());
    if (launches_out) *launches_out = launches;
    return NCDE_OK;
}
