// engine.cuh - engine object: device weights, workspaces, launch bookkeeping, GEMM op builder.
#pragma once
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/ptts_b200.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace ptts {

thread_local char g_last_error[512] = {0};
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------- profiling classes
enum KClass : int {
    KC_GEMV = 0,       // batch-1 bf16 weight-streaming GEMV
    KC_GEMM_FLOW,      // tcgen05 GEMM, batched FlowLM transformer + flow head
    KC_ATTN_DECODE,    // RoPE + KV append + decode attention
    KC_ROWNORM,        // LayerNorm / adaLN / split-K fold
    KC_FLOW_MISC,      // prepare / eos / flow init / flow update
    KC_GEMM_MIMI,      // tcgen05 GEMM, Mimi transformer linears + SEANet implicit GEMM
    KC_MIMI_ATTN,      // banded attention
    KC_MIMI_MISC,      // front end, rope split, final conv
    KC_COUNT
};
static const char *const KCLASS_NAMES[KC_COUNT] = {"gemv_bf16", "gemm_tc_flowlm", "attn_decode", "rownorm",
                                                   "flow_misc", "gemm_tc_mimi", "mimi_win_attn", "mimi_misc"};

struct Prof {
    bool on = false;
    struct Rec { int cls; cudaEvent_t a, b; int rows, n, k, bn; double flops, bytes; };
    std::vector<Rec> recs;
    float ms[KC_COUNT] = {0};
    int64_t n[KC_COUNT] = {0};
    // algorithmic work issued per class while profiling (host-side accounting from launch shapes)
    double flops[KC_COUNT] = {0}, bytes[KC_COUNT] = {0};
    int cur_step = 0;   // decode iteration being launched (profile mode runs without graphs)
    double mean_own_base = 0.0;   // mean prompt rows per utterance already in the own cache
};

// ---------------------------------------------------------------- TMA host API
// The one piece of process-wide state: the driver entry point of cuTensorMapEncodeTiled, resolved once (std::call_once, so
// engines created concurrently from several host threads do not race on it); read-only afterwards.
TmaApi g_tma;
int TmaApi::init() {
    static std::once_flag once;
    std::call_once(once, [this] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn && q == cudaDriverEntryPointSuccess)
            encode = (EncodeTiled)fn;
    });
    PTTS_CHECK(encode != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    return 0;
}

int make_tmap(CUtensorMap *out, const void *base, bool bf16, int k, int rows, int batches, size_t ld, size_t batch_stride,
              int box_rows) {
    if (g_tma.init() != 0) return -1;
    PTTS_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0 && (batch_stride * 2) % 16 == 0,
               "make_tmap: base/strides must be 16-byte aligned");
    PTTS_CHECK(box_rows >= 1 && box_rows <= 256, "make_tmap: box rows %d out of range", box_rows);
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batches};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batches > 1 ? batch_stride : (size_t)rows * ld) * 2};
    cuuint32_t box[3] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_tma.encode(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                              const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (k=%d rows=%d batches=%d ld=%zu)", (int)r, k, rows,
               batches, ld);
    return 0;
}

// ---------------------------------------------------------------- kernel launch (programmatic dependent launch)
// All kernels go out with the programmatic-stream-serialization attribute, so a kernel's launch
// latency, prologue (barrier init, TMEM allocation, tensor-map prefetch) and - for the GEMMs - the
// first weight tiles overlap the tail of its predecessor; each kernel orders itself with pdl_wait()
// (common.cuh). Stream capture turns these into programmatic edges of the step graph.
// PTTS_B200_NO_PDL=1 falls back to plain serialized launches (A/B measurements).
static bool pdl_enabled() {
    static const bool v = getenv("PTTS_B200_NO_PDL") == nullptr;   // initialised once, thread-safe
    return v;
}
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors are picked up by launch_check / cudaGetLastError
}
// The same launch with an L2 access-policy window over [win, win + win_bytes): lines of the window are kept in the persisting
// part of L2 with probability hit_ratio (the rest streams). Used by the batch-1 GEMV path for the weight matrix each launch
// streams; a launch attribute, so it is recorded into the node when the step is captured into a CUDA graph.
template <typename... KArgs, typename... Args>
static void launch_k_window(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const void *win, size_t win_bytes,
                            float hit_ratio, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (win && win_bytes > 0 && hit_ratio > 0.0f) {
        at[na].id = cudaLaunchAttributeAccessPolicyWindow;
        at[na].val.accessPolicyWindow.base_ptr = const_cast<void *>(win);
        at[na].val.accessPolicyWindow.num_bytes = win_bytes;
        at[na].val.accessPolicyWindow.hitRatio = hit_ratio;
        at[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na++;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------- GEMM dispatch
template <int BN, int SPLIT, int EPI = EPI_GENERIC>
static int gemm_launch_t(const GemmOp &op, cudaStream_t st) {
    // the dynamic shared-memory attribute is per device: ptts_b200_create sets it for the engine's device (gemm_set_attrs)
    launch_k(gemm_tc_kernel<BN, SPLIT, EPI>, op.grid, GEMM_THREADS, op.smem_bytes, st, op.mapA, op.mapAlo, op.mapB, op.mapBlo, op.mapOH, op.mapOR, op.p);
    PTTS_CUDA_OK(cudaGetLastError());
    return 0;
}

int gemm_launch(const GemmOp &op, cudaStream_t st) {
    if (op.epi == EPI_CONV && op.split == 0) {
        if (op.bn == 128) return gemm_launch_t<128, 0, EPI_CONV>(op, st);
        if (op.bn == 256) return gemm_launch_t<256, 0, EPI_CONV>(op, st);
    }
    if (op.bn == 256 && op.split == 0) {
        if (op.epi == EPI_RESID) return gemm_launch_t<256, 0, EPI_RESID>(op, st);
        if (op.epi == EPI_GELU) return gemm_launch_t<256, 0, EPI_GELU>(op, st);
        if (op.epi == EPI_ROPE) return gemm_launch_t<256, 0, EPI_ROPE>(op, st);
    }
#define GL(BN_, SP_) if (op.bn == BN_ && op.split == SP_) return gemm_launch_t<BN_, SP_>(op, st);
    GL(32, 0) GL(64, 0) GL(128, 0) GL(256, 0)
    GL(32, 1) GL(64, 1) GL(128, 1) GL(256, 1)
    GL(32, 2) GL(64, 2) GL(128, 2) GL(256, 2)
#undef GL
    set_error("gemm_launch: unsupported tile bn=%d split=%d", op.bn, op.split);
    return -1;
}

// 3D map over a 16-bit OUTPUT tensor [batches][rows][ld] for the TMA-store epilogue: box = 64 columns x 32 rows,
// 128B swizzle (the staging layout of gemm_tc.cuh); rows past `rows` and columns past `cols` are clipped.
int make_store_tmap(CUtensorMap *out, const void *base, bool bf16, int cols, int rows, int batches, size_t ld) {
    if (g_tma.init() != 0) return -1;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_tma.encode(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTTS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed with %d (cols=%d rows=%d batches=%d ld=%zu)", (int)r, cols, rows, batches, ld);
    return 0;
}

static int stage_bytes(int bn, int split) {
    return GEMM_BM * GEMM_BK * 2 * (split == 1 ? 2 : 1) + bn * GEMM_BK * 2 * (split == 2 ? 2 : 1);
}
// one persistent CTA per SM: the whole shared memory (minus barriers) is the TMA ring
int gemm_pick_stages(int bn, int split, int iters, bool) {
    int s = (200 * 1024) / stage_bytes(bn, split);
    if (s > GEMM_MAX_STAGES) s = GEMM_MAX_STAGES;
    if (s < 1) s = 1;
    (void)iters;
    return s;
}

// Description of one contraction; see gemm_tc.cuh for the operand roles.
struct GemmDesc {
    const void *A = nullptr, *Alo = nullptr;   // [a_batches][a_rows][a_k] 16-bit, row stride a_ld
    int a_k = 0, a_rows = 0, a_batches = 1;
    size_t a_ld = 0, a_bstride = 0;
    const void *B = nullptr, *Blo = nullptr;   // [b_rows][b_k] 16-bit, row stride b_ld
    int b_k = 0, b_rows = 0;
    size_t b_ld = 0;
    bool bf16 = true;
    int taps = 1, k_per_tap = 0, shift0 = 0, step = 0, a_k_step = 0, b_k_step = 0;
    int bn = 128, split_k = 1;
    bool two_cta = true;
    bool b_static = true;                      // B is a weight matrix (never written by a kernel)
    int num_sms = 148;                         // SM count of the engine's device (persistent grid size)
    GemmParams epi = {};                       // epilogue fields only
};

static int gemm_build(GemmOp &op, const GemmDesc &d) {
    PTTS_CHECK(d.k_per_tap % GEMM_BK == 0 && d.k_per_tap > 0, "gemm_build: K per tap %d must be a multiple of 64", d.k_per_tap);
    PTTS_CHECK(d.bn == 32 || d.bn == 64 || d.bn == 128 || d.bn == 256, "gemm_build: bad N tile %d", d.bn);
    op = GemmOp{};
    op.split = d.Alo ? 1 : (d.Blo ? 2 : 0);
    op.bn = d.bn;
    if (make_tmap(&op.mapA, d.A, d.bf16, d.a_k, d.a_rows, d.a_batches, d.a_ld, d.a_bstride, GEMM_BM) != 0) return -1;
    if (make_tmap(&op.mapAlo, d.Alo ? d.Alo : d.A, d.bf16, d.a_k, d.a_rows, d.a_batches, d.a_ld, d.a_bstride, GEMM_BM) != 0) return -1;
    if (make_tmap(&op.mapB, d.B, d.bf16, d.b_k, d.b_rows, 1, d.b_ld, 0, d.bn) != 0) return -1;
    if (make_tmap(&op.mapBlo, d.Blo ? d.Blo : d.B, d.bf16, d.b_k, d.b_rows, 1, d.b_ld, 0, d.bn) != 0) return -1;
    GemmParams p = d.epi;
    p.m_tiles_per_batch = (d.a_rows + GEMM_BM - 1) / GEMM_BM;
    p.a_rows = d.a_rows;
    p.n_total = d.b_rows;
    p.kblocks_per_tap = d.k_per_tap / GEMM_BK;
    p.taps = d.taps;
    p.a_row_shift0 = d.shift0;
    p.a_row_step = d.step;
    p.a_k_step = d.a_k_step;
    p.b_k_step = d.b_k_step;
    const int iters = p.taps * p.kblocks_per_tap;
    int split_k = d.split_k < 1 ? 1 : d.split_k;
    if (split_k > iters) split_k = iters;
    const int per = (iters + split_k - 1) / split_k;
    split_k = (iters + per - 1) / per;  // no empty slices
    p.split_k = split_k;
    p.iters_per_split = per;
    p.m_tiles_total = d.a_batches * p.m_tiles_per_batch;
    p.n_tiles = (d.b_rows + d.bn - 1) / d.bn;
    const long long tiles = (long long)p.m_tiles_total * p.n_tiles * split_k;
    PTTS_CHECK(tiles < (1ll << 31), "gemm_build: too many tiles");
    p.total_tiles = (int)tiles;
    p.b_static = d.b_static ? 1 : 0;
    // 16-bit outputs of row-major GEMMs leave through TMA tensor stores (the per-thread 32-byte stores cap the
    // store rate at ~11 B/clk per SM: tools/gemm_conv_probe.py)
    const size_t out_rows = (size_t)d.a_batches * d.a_rows;
    // measured per GEMM (tools/mimi_log.py): a win for the output-heavy convolutions (ConvTranspose with its two outputs
    // -17 %, k=3 convs -8 %), a loss where a residual read shares the epilogue (+9 %) and for the long-K linears (the staging
    // buffers cost ring stages) -> enabled for dual-output and narrow (N tile <= 128) GEMMs without a 16-bit residual
    p.tma_store = (!p.transposed && (p.out_h || p.out_raw) && !p.out_l && !p.rope_q && !p.resid_h && (p.out_raw || d.bn <= 128) &&
                   d.bn % 64 == 0 && p.ldh % 8 == 0 && getenv("PTTS_B200_NO_TMA_STORE") == nullptr) ? 1 : 0;
    if (p.tma_store) {
        if (p.out_h && (((uintptr_t)p.out_h & 15) != 0)) p.tma_store = 0;
        if (p.out_raw && (((uintptr_t)p.out_raw & 15) != 0)) p.tma_store = 0;
    }
    p.stage_stride = 4096 * ((p.out_h ? 1 : 0) + (p.out_raw ? 1 : 0));   // per warp: [32 rows][128 B] per staged tensor
    const int staging = p.tma_store ? GEMM_EPI_WARPS * p.stage_stride : 0;
    int stages = (int)((227 * 1024 - 1024 - 1024 - staging) / stage_bytes(d.bn, op.split));
    if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
    if (stages < 2) { stages = gemm_pick_stages(d.bn, op.split, per, d.two_cta); p.tma_store = 0; }
    p.stages = stages;
    p.stage_off = p.stages * stage_bytes(d.bn, op.split) + 1024;   // barriers live in the 1 KB in between; staging stays 1024-aligned
    memset(&op.mapOH, 0, sizeof op.mapOH); memset(&op.mapOR, 0, sizeof op.mapOR);
    if (p.tma_store) {
        if (p.out_h && make_store_tmap(&op.mapOH, p.out_h, p.out_bf16 != 0, d.b_rows, d.a_rows, d.a_batches, (size_t)p.ldh) != 0) return -1;
        if (p.out_raw && make_store_tmap(&op.mapOR, p.out_raw, p.out_bf16 != 0, d.b_rows, d.a_rows, d.a_batches, (size_t)p.ldh) != 0) return -1;
    }
    (void)out_rows;
    p.idesc = umma_idesc_f16(GEMM_BM, d.bn, d.bf16, d.bf16);
    // the SEANet convolutions take the specialised epilogue (gemm_tc.cuh EPI_CONV)
    op.epi = (p.tma_store && p.out_h && p.bias && !p.scale && !p.resid && !p.resid_h && !p.out_f32 && !p.out_l && !p.rope_q && !p.out_bf16 &&
              p.fast_act && (p.act == ACT_ELU || p.act == ACT_NONE) && op.split == 0 && split_k == 1 && (d.bn == 128 || d.bn == 256) &&
              d.b_rows % 64 == 0 && (p.bias_mask == 0 || (p.bias_mask + 1) % 64 == 0) && getenv("PTTS_B200_NO_EPI_CONV") == nullptr)
                 ? EPI_CONV : EPI_GENERIC;
    // the Mimi transformer linears (fp16 operands, N tile 256, no bias): EPI_RESID / EPI_GELU / EPI_ROPE
    if (op.epi == EPI_GENERIC && !p.transposed && !p.tma_store && op.split == 0 && split_k == 1 && d.bn == 256 && !d.bf16 && !p.bias && !p.resid_h &&
        !p.out_raw && !p.out_l && !p.out_bf16 && p.fast_act && d.b_rows % 16 == 0 && getenv("PTTS_B200_NO_EPI_SPECIAL") == nullptr) {
        if (p.scale && p.resid && p.out_f32 && !p.rope_q && (p.act == ACT_NONE || !p.out_h) && d.b_rows <= 512 && p.ldo % 8 == 0 &&
            (!p.out_h || p.ldh % 16 == 0))
            op.epi = EPI_RESID;   // the scale table takes the (unused) staging area: 2 KB are always allocated behind the barriers
        else if (!p.scale && !p.resid && !p.out_f32 && !p.rope_q && p.out_h && p.act == ACT_GELU_TANH && p.ldh % 16 == 0)
            op.epi = EPI_GELU;
        else if (p.rope_q && !p.scale && !p.resid && !p.out_f32 && !p.out_h && d.b_rows == 1536)
            op.epi = EPI_ROPE;
    }
    op.p = p;
    op.grid = dim3((unsigned)std::min<long long>(tiles, d.num_sms), 1, 1);
    op.smem_bytes = p.stages * stage_bytes(d.bn, op.split) + 1024 + 1024 + (p.tma_store ? GEMM_EPI_WARPS * p.stage_stride : 0);
    if (op.epi == EPI_RESID) op.smem_bytes += 2048;   // layer-scale table at stage_off
    return 0;
}

// ---------------------------------------------------------------- host-side weight helpers
static inline uint16_t h_f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float h_bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint16_t h_f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }
static inline float h_h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

struct HostWeights {
    const ptts_b200_tensor *t;
    int n;
    static bool ends_with(const char *s, const char *suf) {
        size_t a = strlen(s), b = strlen(suf);
        return a >= b && strcmp(s + a - b, suf) == 0;
    }
    // lookup order of the reference: exact, prefixed, suffix (ptts_flowlm.c:107-122, ptts_mimi.c:77-95)
    const ptts_b200_tensor *find(const char *name, const char *const *prefixes) const {
        for (int i = 0; i < n; i++) if (strcmp(t[i].name, name) == 0) return &t[i];
        char buf[512];
        for (int p = 0; prefixes[p]; p++) {
            snprintf(buf, sizeof buf, "%s%s", prefixes[p], name);
            for (int i = 0; i < n; i++) if (strcmp(t[i].name, buf) == 0) return &t[i];
        }
        for (int i = 0; i < n; i++) if (ends_with(t[i].name, name)) return &t[i];
        return nullptr;
    }
    // f32 copy (F32 / F16 / BF16 accepted, ptts_safetensors.c:294-334)
    int get(const char *name, const char *const *prefixes, int64_t numel, std::vector<float> &out, bool optional = false) const {
        const ptts_b200_tensor *e = find(name, prefixes);
        if (!e) {
            if (optional) { out.clear(); return 1; }
            set_error("missing tensor %s", name);
            return -1;
        }
        if (e->numel != numel) {
            set_error("tensor %s has %lld elements, expected %lld", e->name, (long long)e->numel, (long long)numel);
            return -1;
        }
        out.resize((size_t)numel);
        if (e->dtype == 0) memcpy(out.data(), e->data, sizeof(float) * (size_t)numel);
        else if (e->dtype == 1) { const uint16_t *s = (const uint16_t *)e->data; for (int64_t i = 0; i < numel; i++) out[i] = h_h2f(s[i]); }
        else if (e->dtype == 2) { const uint16_t *s = (const uint16_t *)e->data; for (int64_t i = 0; i < numel; i++) out[i] = h_bf2f(s[i]); }
        else { set_error("tensor %s: unsupported dtype %d", e->name, e->dtype); return -1; }
        return 0;
    }
};
static const char *const FLOW_PFX[] = {"flow_lm.", nullptr};
static const char *const MIMI_PFX[] = {"mimi.", "model.", nullptr};

// ---------------------------------------------------------------- engine
struct FlowLayerW { __nv_bfloat16 *inproj, *outproj, *l1, *l2; float *n1w, *n1b, *n2w, *n2b; };
struct FlowBlockW { float *lnw, *lnb, *m0b, *m2b; __nv_bfloat16 *m0w, *m2w; };
struct MimiLayerW { __half *inproj, *outproj, *l1, *l2; float *n1w, *n1b, *n2w, *n2b, *ls1, *ls2; };
struct EngineCaches;   // contraction descriptors + step programs, defined next to their users (ptts_b200.cu)
struct ConvW { __half *w; float *b; int cin, cout, k, stride, n_rows /*GEMM N*/, k_per_tap; bool transposed_conv; };


}  // namespace ptts

struct ptts_b200_engine {
    int device = 0;
    cudaStream_t st = nullptr, st_copy = nullptr;   // compute stream; device->host copies of finished PCM chunks
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_chunk = nullptr, ev_copy = nullptr;
    cudaEvent_t ev_m0 = nullptr, ev_m1 = nullptr;   // around the last launch of the persistent step kernel
    int mega_timed = 0;                             // 1: ev_m0 / ev_m1 have been recorded
    int mimi_st_n = 0, mimi_st_frames = 0;          // latents staged by ptts_b200_mimi_stage
    const float *pcm_on_host = nullptr;             // host buffer the last run already delivered its PCM to
    std::vector<void *> allocs;
    int64_t launches = 0;
    ptts::Prof prof;
    int num_sms = 148;
    // batch-1 decode: share of the GEMV weight stream kept in the persisting part of L2 (0: no access-policy windows) and the
    // largest window a launch may name (cudaDevAttrMaxAccessPolicyWindowSize)
    float l2_hit_ratio = 0.0f;
    size_t l2_max_window = 0;
    ptts::EngineCaches *caches = nullptr;

    // ---- FlowLM weights
    __nv_bfloat16 *embed = nullptr, *in_lin = nullptr;
    float *emb_std = nullptr, *emb_mean = nullptr, *bos_proj = nullptr, *onorm_w = nullptr, *onorm_b = nullptr, *eos_w = nullptr,
          *eos_b = nullptr;
    float h_emb_std[32], h_emb_mean[32];
    ptts::FlowLayerW L[ptts::FL_L];
    __nv_bfloat16 *cond_w = nullptr, *inp_w = nullptr, *ada_w = nullptr, *fin_w = nullptr;
    float *cond_b = nullptr, *inp_b = nullptr, *ada_b = nullptr, *fin_b = nullptr;
    ptts::FlowBlockW FB[ptts::FN_DEPTH];
    struct { float *l0w, *l0b, *l2w, *l2b, *alpha, *freqs; } te[2];
    std::unordered_map<int, float *> cbias;   // lsd_steps -> [steps][512] = cond_b + (ts+tt)/2
    float *rope_cos = nullptr, *rope_sin = nullptr;
    int rope_len = 0;

    // ---- Mimi weights
    float *quant_w = nullptr, *ups_w = nullptr, *dec_out_w = nullptr, *dec_out_b = nullptr;
    float h_dec_out_b = 0.0f;
    ptts::MimiLayerW ML[ptts::MI_L];
    ptts::ConvW dec_in, up[3], res1[3], res2[3];
    // ---- voice-cloning encoder (optional: present when the checkpoint carries the encoder tensors, enc.cuh)
    bool has_encoder = false;
    float *enc_in_w = nullptr, *enc_in_b = nullptr;          // Conv1d 1 -> 64, k 7: [64][7] fp32
    ptts::ConvW eres1[3], eres2[3], edown[3], enc_out, down;  // resblocks (64, 128, 256 channels), strided convs, conv k3, downsampler
    ptts::MimiLayerW EL[ptts::MI_L];                          // encoder_transformer
    __half *spk16 = nullptr;                                  // speaker_proj_weight [1024, 512]

    // ---- FlowLM run state (capacity-sized)
    int capB = 0, capS = 0, capTok = 0, capF = 0, capCond = 0;
    int curB = 0;
    ptts::FlowCtrl *ctrl = nullptr;
    int *d_tokens = nullptr, *d_tok_len = nullptr, *d_max_frames = nullptr, *d_eos_after = nullptr;
    float *d_noise = nullptr, *d_latent = nullptr, *d_out_latents = nullptr;
    int *d_eos_step = nullptr, *d_done = nullptr, *d_used = nullptr;
    float *d_first_eos = nullptr, *d_first_cond = nullptr, *d_first_flow = nullptr;
    float *x = nullptr, *qkv = nullptr, *att = nullptr, *ffh = nullptr, *normed = nullptr;
    uint16_t *xn_h = nullptr, *xn_l = nullptr, *att_h = nullptr, *att_l = nullptr, *ffh_h = nullptr, *ffh_l = nullptr,
             *normed_h = nullptr, *normed_l = nullptr;
    float *partA = nullptr, *partB = nullptr;     // split-K partials [4][B][1024]
    float *fx = nullptr, *ada = nullptr, *ft = nullptr, *flow = nullptr, *sc = nullptr, *fg = nullptr;
    uint16_t *sc_h = nullptr, *sc_l = nullptr, *fh_h = nullptr, *fh_l = nullptr, *fg_h = nullptr, *fg_l = nullptr;
    float *pk = nullptr, *pv = nullptr;           // prefix KV [L][H][capCond][64]
    __half *pk16 = nullptr, *pv16 = nullptr;      // fp16 copy (made once per voice)
    __half *ok = nullptr, *ov = nullptr;          // own KV, fp16 [L][H][capB][capS][64]
    float *d_cond = nullptr;                      // voice rows on device [capCond,1024]
    // batched prefill workspace (capRows packed rows)
    int capRows = 0;
    int *d_row_b = nullptr, *d_row_i = nullptr, *d_row_start = nullptr, *d_own_base = nullptr;
    float *pf_x = nullptr, *pf_qkv = nullptr;
    uint16_t *pf_xn_h = nullptr, *pf_xn_l = nullptr, *pf_att_h = nullptr, *pf_att_l = nullptr, *pf_ffh_h = nullptr, *pf_ffh_l = nullptr;
    int st_rows = 0, st_maxq = 0, st_cond = 0;
    std::vector<float> voice_host;                // copy of the cached voice prefix
    int voice_len = -1;
    // step graphs keyed by (mode, B, lsd_steps, precise); `gen` bumps whenever device pointers change
    std::unordered_map<uint64_t, cudaGraphExec_t> graphs;
    std::unordered_map<uint64_t, int64_t> graph_kernels;
    uint64_t gen = 0;
    void invalidate_graphs() {
        for (auto &g : graphs) cudaGraphExecDestroy(g.second);
        graphs.clear();
        graph_kernels.clear();
        gen++;
    }
    // staged batch
    int st_n = 0, st_steps = 0, st_capF = 0, st_lsd = 1, st_precise = 1, st_eos = 0;
    std::vector<int> st_maxf;
    unsigned long long *mega_ctl = nullptr;   // persistent step kernel: [0] grid barrier, [1] done count, [8..] phase timers
    int stream_pos = -1;                      // streaming entry: frames already handed out (-1: no stream open)
    // pinned staging
    void *pin = nullptr;
    size_t pin_bytes = 0;
    // one-shot batched call: the per-utterance noise (data-independent, generated on host threads) is produced and uploaded AFTER
    // the prompt prefill has been launched, so the host work runs under the prefill's device time
    float *pin_noise = nullptr;
    size_t pin_noise_cap = 0;
    struct { int n = 0, cap_f = 0; float temp = 0.0f, clamp = 0.0f; std::vector<int64_t> seeds; std::vector<int> frames; } pend_noise;

    // ---- Mimi workspace
    void *mimi_ws = nullptr;
    size_t mimi_ws_bytes = 0;
    float *d_mimi_lat = nullptr, *d_pcm = nullptr;
    size_t mimi_lat_cap = 0, pcm_cap = 0;
    // streaming decode state of the open stream (mimi.cuh): q scratch and per-layer K / V caches [8][ms_cap][64] fp16,
    // transformer outputs [ms_cap][512] fp16 (source of the SEANet conv tails)
    __half *ms_q = nullptr, *ms_xn = nullptr, *ms_k[ptts::MI_L] = {nullptr, nullptr}, *ms_v[ptts::MI_L] = {nullptr, nullptr};
    int ms_cap = 0;

    template <typename T>
    int dalloc(T **p, size_t count) {
        void *q = nullptr;
        cudaError_t r = cudaMalloc(&q, count * sizeof(T) < 16 ? 16 : count * sizeof(T));
        if (r != cudaSuccess) {
            ptts::set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(r));
            return -1;
        }
        allocs.push_back(q);
        *p = (T *)q;
        return 0;
    }
    void dfree(void *p) {
        if (!p) return;
        for (size_t i = 0; i < allocs.size(); i++)
            if (allocs[i] == p) { allocs[i] = allocs.back(); allocs.pop_back(); break; }
        cudaFree(p);
    }
    int up_f32(float **dst, const std::vector<float> &v) {
        if (dalloc(dst, v.size()) != 0) return -1;
        PTTS_CUDA_OK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    }
    int up_bf16(__nv_bfloat16 **dst, const std::vector<float> &v) {
        std::vector<uint16_t> h(v.size());
        for (size_t i = 0; i < v.size(); i++) h[i] = ptts::h_f2bf(v[i]);
        if (dalloc(dst, v.size()) != 0) return -1;
        PTTS_CUDA_OK(cudaMemcpy(*dst, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        return 0;
    }
    int up_f16(__half **dst, const std::vector<float> &v) {
        std::vector<uint16_t> h(v.size());
        for (size_t i = 0; i < v.size(); i++) h[i] = ptts::h_f2h(v[i]);
        if (dalloc(dst, v.size()) != 0) return -1;
        PTTS_CUDA_OK(cudaMemcpy(*dst, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        return 0;
    }
};

namespace ptts {
typedef ptts_b200_engine Engine;

// Launch bookkeeping: count every launch; in profile mode bracket it with events.
struct LaunchScope {
    Engine *e;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(Engine *e_, int cls_) : e(e_), cls(cls_) {
        e->launches++;
        if (e->prof.on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, e->st);
        }
    }
    ~LaunchScope() {
        if (e->prof.on) {
            cudaEventRecord(b, e->st);
            e->prof.recs.push_back({cls, a, b, 0, 0, 0, 0, 0.0, 0.0});
        }
    }
};

static int launch_check(const char *what) {
    cudaError_t r = cudaGetLastError();
    if (r != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(r));
        return -1;
    }
    return 0;
}

static int run_gemm(Engine *e, const GemmOp &op, int cls) {
    if (e->prof.on) {
        // algorithmic work: every A row against every B row over the full K; operands and outputs once
        const double rows = (double)(op.p.m_tiles_total / op.p.m_tiles_per_batch) * op.p.a_rows, n = op.p.n_total;
        const double k = (double)op.p.taps * op.p.kblocks_per_tap * GEMM_BK;
        const double a_el = op.p.a_k_step == 0 && op.p.taps > 1 ? rows * op.p.kblocks_per_tap * GEMM_BK : rows * k;
        e->prof.flops[cls] += 2.0 * rows * n * k;
        e->prof.bytes[cls] += 2.0 * (a_el * (op.split == 1 ? 2 : 1) + n * k * (op.split == 2 ? 2 : 1)) +
                              rows * n * ((op.p.out_f32 ? 4.0 : 0.0) + (op.p.resid ? 4.0 : 0.0) + (op.p.out_h ? 2.0 : 0.0) + (op.p.resid_h ? 2.0 : 0.0) + (op.p.out_raw ? 2.0 : 0.0) +
                                          (op.p.out_l ? 2.0 : 0.0));
    }
    int rc;
    {
        LaunchScope ls(e, cls);
        rc = gemm_launch(op, e->st);
    }
    if (e->prof.on && !e->prof.recs.empty()) {   // shape tag for the per-launch log (PTTS_B200_GEMM_LOG=1)
        Prof::Rec &r = e->prof.recs.back();
        r.rows = (op.p.m_tiles_total / op.p.m_tiles_per_batch) * op.p.a_rows; r.n = op.p.n_total;
        r.k = op.p.taps * op.p.kblocks_per_tap * GEMM_BK; r.bn = op.bn;
        r.flops = 2.0 * r.rows * (double)r.n * r.k;
    }
    return rc;
}

}  // namespace ptts
