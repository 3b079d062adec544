// Actor-critic MLP kernels on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as the SIMT kernels in mlp.cu (ts_ppo_grad / ts_critic_forward / ts_actor_logp);
// reference code replaced: modelfree/ppo.py:157-161,179-211, modelfree/a2c.py:123-126,
// algorithm_base.py:497.
//
// Numerics: every GEMM is an fp32-faithful product built from bf16 tensor-core MMAs with fp32
// accumulation in TMEM: each operand element x is stored as three bf16 pieces b0 + b1 + b2 = x
// (24 significant bits) and the six partial products of weight >= 2^-16 are accumulated
// (umma::gemm_bf16x3).  A single-pass bf16/tf32 MMA would be 6x / 3x cheaper but is ~1e-3 off
// the reference's fp32 results, which breaks the 1e-5 parity bar on v_s / returns / advantages.
//
// Layout: one CTA = one tile of 128 transitions = the 128 TMEM lanes.  Every operand matrix
// (activations X, H1, H2, gradients, weights) lives in shared memory in the blocked no-swizzle
// layout of umma.cuh (8-row x 16-byte core matrices), ONE copy per matrix: the same bytes are
// consumed K-major by the forward / input-gradient GEMMs and MN-major (reduction over the 128
// rows) by the weight-gradient GEMMs.  Activations never leave the SM between forward and backward:
//   X -(W1)-> D1 -tanh-> H1 -(W2)-> D2 -tanh-> H2 -(W3)-> D3 -> loss -> dOut
//   dW3 = H2^T dOut ; dZ2 = (dOut W3) (1-H2^2) [overwrites H2] ; dW2 = dZ2^T H1 ; db2 = dZ2^T 1 ;
//   dH1 = dZ2 W2 ; dZ1 = dH1 (1-H1^2) [overwrites H1] ; dW1 = dZ1^T X ; db1 = dZ1^T 1
// Bias gradients come out of the tensor core too (B operand = a 128 x 8 block of ones).
// MMAs are issued by one thread, completion is signalled through an mbarrier (tcgen05.commit);
// the 8 warps do the TMEM -> register epilogues (tanh, loss, splits) and the gradient REDs.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "ppo_math.cuh"
#include "umma.cuh"

namespace {

constexpr int H = 64;
constexpr int kRows = 128;
constexpr int kThreads = 512;        // 16 warps: 4 per TMEM sub-partition, 16 accumulator columns per thread
constexpr int kCols = 16;            // columns of a 64-wide layer owned by one thread in the epilogues
constexpr int kMaxAct = 16;
constexpr int NO = 16;            // padded head width (N of the head GEMM, columns of dOut)

// TMEM column map (fp32 accumulators)
// dW2 / dW1 carry one extra 8-column block: the bias gradient (B operand extended by a block of ones)
constexpr uint32_t cD1 = 0, cD2 = 64, cD3 = 128, cDH1 = 192, cDW2 = 256 /* 72 */, cDW3 = 328 /* 16 */, cDW1 = 344 /* <= 40 */;
// TS-mode A operand (bf16x3 pieces of the CURRENT activation / gradient tile: H1, then H2, then dZ2), 3 x 32 packed columns
constexpr uint32_t cT = 384, kTPart = 32;
constexpr uint32_t kTmemCols = 512;

// Optional phase timeline (DIAGNOSTICS build only, -DTS_B200_DIAGNOSTICS -> libts_b200_diag.so): when enabled, one CTA /
// thread 0 stores %globaltimer at each phase boundary; read back with ts_tc_timeline().  Compiled out of the product library.
#ifdef TS_B200_DIAGNOSTICS
__device__ unsigned long long g_tc_timeline[32];
__device__ int g_tc_timeline_on = 0;
__device__ int g_tc_timeline_gate = 1;      // written and read by CTA 0 / thread 0 only: which step is recorded
__device__ __forceinline__ void tstamp(int slot) {
    if (g_tc_timeline_on && (int)blockIdx.x == g_tc_timeline_on - 1 && threadIdx.x == 0 && g_tc_timeline_gate) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tc_timeline[slot] = t;
    }
}
#else
__device__ __forceinline__ void tstamp(int) {}
#endif


// ---- bf16x3 operand matrices in shared memory --------------------------------------------------
struct Mat {
    uint32_t base;   // shared address of piece 0
    uint32_t part;   // bytes between pieces
    uint32_t RS;     // bytes between 8-row groups (= cols/8 * 128)
};
__host__ __device__ inline uint32_t mat_bytes(int rows, int cols) { return (uint32_t)rows * cols * 2u; }
__device__ __forceinline__ uint32_t moff(uint32_t r, uint32_t c, uint32_t RS) {
    return (r >> 3) * RS + (c >> 3) * 128u + (r & 7u) * 16u + (c & 7u) * 2u;
}

// x = b0 + b1 + b2 with three bf16 pieces obtained by TRUNCATION (x & 0xffff0000), each capturing 8
// more significant bits; the residuals are exact in fp32.  4 ALU ops per element + 3 PRMT per pair.
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
    const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
    w0 = __byte_perm(u0, u1, 0x7632);                       // high halves of (x0, x1)
    const float r0 = x0 - __uint_as_float(u0 & 0xffff0000u), r1 = x1 - __uint_as_float(u1 & 0xffff0000u);
    const uint32_t v0 = __float_as_uint(r0), v1 = __float_as_uint(r1);
    w1 = __byte_perm(v0, v1, 0x7632);
    const float s0 = r0 - __uint_as_float(v0 & 0xffff0000u), s1 = r1 - __uint_as_float(v1 & 0xffff0000u);
    w2 = __byte_perm(__float_as_uint(s0), __float_as_uint(s1), 0x7632);
}
// 8 consecutive columns (one 16-byte chunk) of row r
__device__ __forceinline__ void store_chunk8(uint8_t* sm0, const Mat& m, uint32_t r, uint32_t c0, const float* v) {
    uint32_t w0[4], w1[4], w2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3_pair(v[2 * j], v[2 * j + 1], w0[j], w1[j], w2[j]);
    uint8_t* p = sm0 + (m.base + moff(r, c0, m.RS));
    *reinterpret_cast<uint4*>(p) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    *reinterpret_cast<uint4*>(p + m.part) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    *reinterpret_cast<uint4*>(p + 2 * m.part) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
}

// 16 consecutive columns [c0, c0 + 16) of row r: bf16x3 split once, stored to the shared-memory operand (consumed
// MN-major by the weight-gradient GEMMs) AND to the thread's TMEM lane (A operand of the next TS-mode GEMM:
// packed columns t_col + c0 / 2 of each piece).  t_lane_col = 0xffffffff: shared memory only.
__device__ __forceinline__ void store_row16(uint8_t* sm0, const Mat& m, uint32_t r, uint32_t c0, const float* v,
                                            uint32_t t_lane_col) {
    uint32_t w0[8], w1[8], w2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3_pair(v[2 * j], v[2 * j + 1], w0[j], w1[j], w2[j]);
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
        uint8_t* p = sm0 + (m.base + moff(r, c0 + 8 * hlf, m.RS));
        *reinterpret_cast<uint4*>(p) = make_uint4(w0[4 * hlf], w0[4 * hlf + 1], w0[4 * hlf + 2], w0[4 * hlf + 3]);
        *reinterpret_cast<uint4*>(p + m.part) = make_uint4(w1[4 * hlf], w1[4 * hlf + 1], w1[4 * hlf + 2], w1[4 * hlf + 3]);
        *reinterpret_cast<uint4*>(p + 2 * m.part) = make_uint4(w2[4 * hlf], w2[4 * hlf + 1], w2[4 * hlf + 2], w2[4 * hlf + 3]);
    }
    if (t_lane_col != 0xffffffffu) {
        umma::tmem_st8(t_lane_col + (c0 >> 1), w0);
        umma::tmem_st8(t_lane_col + (c0 >> 1) + kTPart, w1);
        umma::tmem_st8(t_lane_col + (c0 >> 1) + 2 * kTPart, w2);
    }
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) from two MUFU ops; absolute error ~1e-7 (the hidden activations
// feed 64-term dot products, so absolute -- not relative -- accuracy near 0 is what matters)
__device__ __forceinline__ float tanh_mufu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// D[M x N] = A * B (operand usage K-major / MN-major per flag).  WARP-LEVEL: call from all lanes of
// one warp; K = 16 * KSTEPS.
template <int KSTEPS, bool FULL = true>
__device__ __forceinline__ void gemm(uint32_t d_tmem, int M, int N, const Mat& A, int a_mn, const Mat& B, int b_mn) {
    const uint32_t a_lbo = a_mn ? A.RS : 128u, a_sbo = a_mn ? 128u : A.RS, a_step = a_mn ? 2u * A.RS : 256u;
    const uint32_t b_lbo = b_mn ? B.RS : 128u, b_sbo = b_mn ? 128u : B.RS, b_step = b_mn ? 2u * B.RS : 256u;
    umma::gemm_bf16x3_warp<KSTEPS, 3, 3, FULL>(d_tmem, A.base, A.part, a_lbo, a_sbo, a_step, B.base, B.part, b_lbo, b_sbo, b_step,
                                               umma::idesc_bf16(M, N, a_mn, b_mn));
}
// weight-gradient GEMMs: three-product scheme unless TS_B200_WGRAD_FULL is defined at build time
#ifdef TS_B200_WGRAD_FULL
constexpr bool kWgradFull = true;
#else
constexpr bool kWgradFull = false;
#endif
// TS mode: A = the bf16x3 pieces at TMEM columns cT (written by the preceding epilogue), M = 128, K = 64
__device__ __forceinline__ void gemm_ts(uint32_t tmem, uint32_t d_col, int N, const Mat& B, int b_mn) {
    const uint32_t b_lbo = b_mn ? B.RS : 128u, b_sbo = b_mn ? 128u : B.RS, b_step = b_mn ? 2u * B.RS : 256u;
    umma::gemm_bf16x3_ts_warp<H / 16>(tmem + d_col, tmem + cT, kTPart, B.base, B.part, b_lbo, b_sbo, b_step,
                                      umma::idesc_bf16(128, N, 0, b_mn));
}
// runtime K in {16, 32} (padded observation width)
__device__ __forceinline__ void gemm_kx(uint32_t d_tmem, int M, int N, const Mat& A, int a_mn, const Mat& B, int b_mn, int K) {
    if (K == 16) gemm<1>(d_tmem, M, N, A, a_mn, B, b_mn); else gemm<2>(d_tmem, M, N, A, a_mn, B, b_mn);
}
struct Smem {   // byte offsets from the dynamic shared memory base (all multiples of 128)
    int KXP;
    Mat X, H1, H2, DO, W1, W2, W3;
    uint32_t w3f, b1, b2, b3, ls, dof, rowv, red, act;
    uint32_t wblk, wblk_bytes;   // the "weight block" W1 | W2 | W3 | w3f | b1 | b2 | b3 | ls: one contiguous range, the unit
                                 // of the pre-split weight image in global memory (one bulk copy per network)
    uint32_t total;
};
__host__ __device__ inline Smem make_smem(int obs_dim, uint32_t sbase) {
    Smem s;
    s.KXP = (obs_dim + 15) & ~15;
    uint32_t o = 0;
    auto mat = [&](Mat& m, int rows, int cols) {
        m.base = sbase + o; m.part = mat_bytes(rows, cols); m.RS = (uint32_t)(cols / 8) * 128u; o += 3u * m.part;
    };
    // X and H1 carry one extra 8-column chunk of ones (piece 0 = 1.0, pieces 1, 2 = 0): as the B operand of the
    // weight-gradient GEMMs it yields the bias gradient as 8 more accumulator columns instead of a GEMM of its own
    mat(s.X, kRows, s.KXP + 8);
    mat(s.H1, kRows, H + 8);
    mat(s.H2, kRows, H);
    mat(s.DO, kRows, NO);
    s.wblk = o;
    mat(s.W1, H, s.KXP);
    mat(s.W2, H, H);
    mat(s.W3, NO, H);
    s.w3f = o;  o += kMaxAct * H * 4;      // natural fp32 W3 [a][k] for the SIMT K=act GEMM
    s.b1 = o;   o += H * 4;
    s.b2 = o;   o += H * 4;
    s.b3 = o;   o += kMaxAct * 4;
    s.ls = o;   o += kMaxAct * 4;
    s.wblk_bytes = o - s.wblk;             // multiple of 128
    s.dof = o;  o += kRows * kMaxAct * 4;  // dOut in fp32 [a][r] (rows on consecutive banks)
    s.act = o;  o += kRows * kMaxAct * 4;  // actions of the tile
    s.rowv = o; o += 4 * kRows * 4;        // adv, ret, logp_old, v_s
    s.red = o;  o += 256 * 4;              // [0,12) per-warp sums, [64,80) 1/var, [80,96) log sigma + log sqrt(2 pi),
                                           // [128,256) per-warp column sums of (dmu, dlogstd)
    s.total = o;
    return s;
}

struct NetG { int64_t w1, b1, w2, b2, w3, b3, ls; };

// Stage a [rows x cols] fp32 matrix (row pointer by functor, cols padded with zeros up to ncols_pad)
// into a bf16x3 operand: one thread = one (row, 8-column chunk) -> 8 loads in flight, 3 STS.128, no
// integer division (chunks per row is a power of two).
template <class RowPtrF>
__device__ __forceinline__ void stage_chunks(uint8_t* sm0, const Mat& M, int rows, int cols, int cols_pad,
                                             RowPtrF&& rowptr) {
    const int nch = cols_pad >> 3;                 // 2, 4 or 8
    const int sh = nch == 8 ? 3 : (nch == 4 ? 2 : 1);
    for (int task = threadIdx.x; task < rows * nch; task += kThreads) {
        const int r = task >> sh, ch = task & (nch - 1);
        const float* src = rowptr(r);              // nullptr -> zero row
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * ch + j;
            v[j] = (src != nullptr && k < cols) ? __ldg(src + k) : 0.0f;
        }
        store_chunk8(sm0, M, r, 8 * ch, v);
    }
}

// Split form of stage_chunks for matrices with at most one task per thread: issue the 8 loads of the
// thread's chunk now (chunk_load), convert + store later (chunk_store), so that the loads of several
// matrices are in flight together.
// COHERENT: read through L2 (ld.global.cg) -- required for the parameters, which other CTAs rewrite
// between the steps of the persistent epoch kernel (the read-only / L1 path could return stale lines).
template <bool COHERENT = false, class RowPtrF>
__device__ __forceinline__ void chunk_load(int rows, int cols, int cols_pad, RowPtrF&& rowptr, float (&v)[8]) {
    const int nch = cols_pad >> 3;
    const int sh = nch == 8 ? 3 : (nch == 4 ? 2 : 1);
    const int task = threadIdx.x;
    const int r = task >> sh, ch = task & (nch - 1);
    const float* src = task < rows * nch ? rowptr(r) : (const float*)nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = 8 * ch + j;
        v[j] = (src != nullptr && k < cols) ? (COHERENT ? __ldcg(src + k) : __ldg(src + k)) : 0.0f;
    }
}
__device__ __forceinline__ void chunk_store(uint8_t* sm0, const Mat& M, int rows, int cols_pad, const float (&v)[8]) {
    const int nch = cols_pad >> 3;
    const int sh = nch == 8 ? 3 : (nch == 4 ? 2 : 1);
    const int task = threadIdx.x;
    if (task < rows * nch) store_chunk8(sm0, M, task >> sh, 8 * (task & (nch - 1)), v);
}

// stage one network's weights: bf16x3 blocked copies for the tensor core + fp32 side copies
__device__ void stage_weights(uint8_t* sm, uint8_t* sm0, const Smem& S, const float* params, const NetG& g,
                              int obs_dim, int out_dim) {
    const int tid = threadIdx.x;
    float* w3f = reinterpret_cast<float*>(sm + S.w3f);
    {   // <= one chunk per thread and matrix (H = 64, KXP <= 32, 512 threads): all loads first
        float v1[8], v2[8], v3[8], vf[2];
        chunk_load<true>(H, obs_dim, S.KXP, [&](int o) { return params + g.w1 + (int64_t)o * obs_dim; }, v1);
        chunk_load<true>(H, H, H, [&](int o) { return params + g.w2 + (int64_t)o * H; }, v2);
        chunk_load<true>(NO, H, H, [&](int a) { return a < out_dim ? params + g.w3 + (int64_t)a * H : (const float*)nullptr; }, v3);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = tid + u * kThreads;
            vf[u] = (e >> 6) < out_dim ? __ldcg(params + g.w3 + e) : 0.0f;
        }
        chunk_store(sm0, S.W1, H, S.KXP, v1);
        chunk_store(sm0, S.W2, H, H, v2);
        chunk_store(sm0, S.W3, NO, H, v3);
        w3f[tid] = vf[0]; w3f[tid + kThreads] = vf[1];
    }
    float* b1 = reinterpret_cast<float*>(sm + S.b1);
    float* b2 = reinterpret_cast<float*>(sm + S.b2);
    float* b3 = reinterpret_cast<float*>(sm + S.b3);
    float* ls = reinterpret_cast<float*>(sm + S.ls);
    for (int e = tid; e < H; e += kThreads) { b1[e] = __ldcg(params + g.b1 + e); b2[e] = __ldcg(params + g.b2 + e); }
    if (tid < kMaxAct) {
        b3[tid] = tid < out_dim ? __ldcg(params + g.b3 + tid) : 0.0f;
        const float l = (tid < out_dim && g.ls >= 0) ? __ldcg(params + g.ls + tid) : 0.0f;
        ls[tid] = l;
        // per-dimension constants of the diagonal Gaussian, once per CTA instead of once per row
        float* gs = reinterpret_cast<float*>(sm + S.red) + 64;
        const float sigma = expf(l);
        gs[tid] = 1.0f / (sigma * sigma);
        gs[16 + tid] = logf(sigma) + 0.9189385332046727f;
    }
}

// ---- pre-split weight image --------------------------------------------------------------------------
// Global-memory copy of both networks' weight blocks in exactly the shared-memory layout (bf16x3 blocked
// operands + fp32 side copies; block 0 = critic, block 1 = actor), so that staging a network is ONE
// cp.async.bulk instead of ~13 k scattered loads + splits per CTA and step.  Built by
// weight_image_build_kernel; the Adam phase of the epoch kernel updates the entries of the parameters it
// rewrites.  Padding (columns >= obs_dim, head rows >= out_dim) is zero and never touched.
__device__ __forceinline__ void img_put(uint8_t* blk, uint32_t rel, uint32_t part, uint32_t RS, uint32_t r, uint32_t c, float x) {
    uint32_t w0, w1, w2;
    split3_pair(x, 0.0f, w0, w1, w2);
    uint8_t* p = blk + rel + moff(r, c, RS);
    *reinterpret_cast<uint16_t*>(p) = (uint16_t)w0;
    *reinterpret_cast<uint16_t*>(p + part) = (uint16_t)w1;
    *reinterpret_cast<uint16_t*>(p + 2 * part) = (uint16_t)w2;
}
__device__ __forceinline__ bool img_scatter_net(const Smem& S, uint32_t sbase, const NetG& g, int obs_dim, int out_dim,
                                                int64_t i, float x, uint8_t* blk) {
    const uint32_t w0 = sbase + S.wblk;       // Mat bases are shared addresses; the image uses block-relative offsets
    int64_t o;
    if ((o = i - g.w1) >= 0 && o < (int64_t)H * obs_dim) {
        const uint32_t r = (uint32_t)o / (uint32_t)obs_dim, c = (uint32_t)o - r * (uint32_t)obs_dim;
        img_put(blk, S.W1.base - w0, S.W1.part, S.W1.RS, r, c, x);
        return true;
    }
    if ((o = i - g.w2) >= 0 && o < H * H) { img_put(blk, S.W2.base - w0, S.W2.part, S.W2.RS, (uint32_t)o >> 6, (uint32_t)o & 63u, x); return true; }
    if ((o = i - g.w3) >= 0 && o < (int64_t)out_dim * H) {
        img_put(blk, S.W3.base - w0, S.W3.part, S.W3.RS, (uint32_t)o >> 6, (uint32_t)o & 63u, x);
        reinterpret_cast<float*>(blk + (S.w3f - S.wblk))[o] = x;
        return true;
    }
    if ((o = i - g.b1) >= 0 && o < H) { reinterpret_cast<float*>(blk + (S.b1 - S.wblk))[o] = x; return true; }
    if ((o = i - g.b2) >= 0 && o < H) { reinterpret_cast<float*>(blk + (S.b2 - S.wblk))[o] = x; return true; }
    if ((o = i - g.b3) >= 0 && o < out_dim) { reinterpret_cast<float*>(blk + (S.b3 - S.wblk))[o] = x; return true; }
    if (g.ls >= 0 && (o = i - g.ls) >= 0 && o < out_dim) { reinterpret_cast<float*>(blk + (S.ls - S.wblk))[o] = x; return true; }
    return false;
}
__device__ __forceinline__ void img_scatter(const ts_actor_critic_desc& d, const Smem& S, uint32_t sbase, int64_t i, float x,
                                            uint8_t* wimg) {
    const NetG gc{d.c_w1, d.c_b1, d.c_w2, d.c_b2, d.c_w3, d.c_b3, -1};
    if (img_scatter_net(S, sbase, gc, d.obs_dim, 1, i, x, wimg)) return;
    const NetG ga{d.a_w1, d.a_b1, d.a_w2, d.a_b2, d.a_w3, d.a_b3, d.a_logstd};
    img_scatter_net(S, sbase, ga, d.obs_dim, d.act_dim, i, x, wimg + S.wblk_bytes);
}
__global__ void weight_image_build_kernel(const float* __restrict__ params, const ts_actor_critic_desc d, uint8_t* __restrict__ wimg) {
    const Smem S = make_smem(d.obs_dim, 0u);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_params; i += (int64_t)gridDim.x * blockDim.x)
        img_scatter(d, S, 0u, i, params[i], wimg);
}

struct Pipe {   // MMA issue / completion handshake
    uint64_t* bar;
    uint32_t phase;
    // all threads: make smem writes + tcgen05.ld's visible; then WARP 0 (all lanes, warp-uniform
    // control flow) runs `f`, whose MMAs are issued by one elected lane, and commits.
    template <class F>
    __device__ __forceinline__ void issue(F&& f) {
        umma::tmem_wait_st();
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
        if (warp_u == 0) {
            umma::fence_after_sync();
            f();
            if (umma::elect_one()) umma::mma_commit(bar);
            __syncwarp();
        }
    }
    // all threads: the MMAs of the last issue() have completed (work that does not touch their operands or
    // accumulators may run between issue() and wait())
    __device__ __forceinline__ void wait() {
        umma::mbar_wait(bar, phase);
        phase ^= 1u;
        umma::fence_after_sync();
    }
    template <class F>
    __device__ __forceinline__ void run(F&& f) { issue(f); wait(); }
};

// Epilogue thread map: warp w -> TMEM sub-partition q = w & 3 (rows 32q + lane), column group
// cq = w >> 2 -> columns [16 cq, 16 cq + 16) of a 64-wide accumulator.
// TMEM -> h = tanh(x + bias) -> bf16x3 rows of OUT; h stays in registers for the backward pass
__device__ __forceinline__ void epi_tanh(uint8_t* sm0, const Mat& OUT, uint32_t tmem, uint32_t col,
                                         const float* __restrict__ bias, float (&h)[kCols]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = 32u * (warp & 3) + lane, c0 = (uint32_t)kCols * (warp >> 2);
    umma::tmem_ld16(tmem + ((32u * (warp & 3)) << 16) + col + c0, h);
#pragma unroll
    for (int j = 0; j < kCols; ++j) h[j] = tanh_mufu(h[j] + bias[c0 + j]);
    store_row16(sm0, OUT, r, c0, h, tmem + ((32u * (warp & 3)) << 16) + cT);
}
// TMEM dH -> dZ = dH * (1 - h^2) written over ACT (h from registers)
__device__ __forceinline__ void epi_dtanh(uint8_t* sm0, const Mat& ACT, uint32_t tmem, uint32_t col,
                                          const float (&h)[kCols]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = 32u * (warp & 3) + lane, c0 = (uint32_t)kCols * (warp >> 2);
    float v[kCols];
    umma::tmem_ld16(tmem + ((32u * (warp & 3)) << 16) + col + c0, v);
#pragma unroll
    for (int j = 0; j < kCols; ++j) v[j] = v[j] * fmaf(-h[j], h[j], 1.0f);
    store_chunk8(sm0, ACT, r, c0, v);
    store_chunk8(sm0, ACT, r, c0 + 8, v + 8);
}
// dZ2 = (dOut W3) * (1 - H2^2) (K = out_dim is tiny: SIMT).  Split in two so that the arithmetic overlaps the dW3 MMA
// that is still reading H2: compute into registers, then (after the MMA completed) store over H2 and into TMEM.
__device__ __forceinline__ void head_input_grad_compute(uint8_t* sm, const Smem& S, int out_dim, const float (&h)[kCols],
                                                        float (&acc)[kCols]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = 32u * (warp & 3) + lane, c0 = (uint32_t)kCols * (warp >> 2);
    const float* dof = reinterpret_cast<const float*>(sm + S.dof) + r;
    const float* w3f = reinterpret_cast<const float*>(sm + S.w3f);
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[j] = 0.0f;
    for (int a = 0; a < out_dim; ++a) {
        const float dv = dof[a * kRows];
#pragma unroll
        for (int j = 0; j < kCols; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w3f + a * H + c0 + j);
            acc[j] = fmaf(dv, w.x, acc[j]); acc[j + 1] = fmaf(dv, w.y, acc[j + 1]);
            acc[j + 2] = fmaf(dv, w.z, acc[j + 2]); acc[j + 3] = fmaf(dv, w.w, acc[j + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[j] = acc[j] * fmaf(-h[j], h[j], 1.0f);
}
__device__ __forceinline__ void head_input_grad_store(uint8_t* sm0, const Smem& S, uint32_t tmem, const float (&acc)[kCols]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = 32u * (warp & 3) + lane, c0 = (uint32_t)kCols * (warp >> 2);
    store_row16(sm0, S.H2, r, c0, acc, tmem + ((32u * (warp & 3)) << 16) + cT);
}
// write one row of dOut: fp32 side copy + bf16x3 operand (columns >= out_dim are zero)
__device__ __forceinline__ void write_dout_row(uint8_t* sm, uint8_t* sm0, const Smem& S, uint32_t r, const float* dv) {
    float* dof = reinterpret_cast<float*>(sm + S.dof) + r;
#pragma unroll
    for (int a = 0; a < kMaxAct; ++a) dof[a * kRows] = dv[a];
    store_chunk8(sm0, S.DO, r, 0, dv);
    store_chunk8(sm0, S.DO, r, 8, dv + 8);
}

// The partial row is private to the CTA: the first tile of a CTA stores, later tiles read-modify-write.
__device__ __forceinline__ void out_acc(float* p, float v, bool first) { *p = first ? v : *p + v; }

// Weight-gradient accumulators (TMEM, M = 64: row o = 16 q + lane for lane < 16) -> the CTA's partial
// gradient row.  A lane owns a ROW of an accumulator, so direct stores would touch one 32-byte sector
// per value (measured: 4-8 us per net).  The tile is transposed through shared memory instead (the H2
// operand is dead by then; padded leading dimensions keep both sides bank-conflict free) and
// written out with consecutive threads on consecutive addresses.
constexpr int kLdW2 = H + 1, kLdW1 = 33, kScrW2 = 0, kScrW1 = kScrW2 + H * kLdW2, kScrW3 = kScrW1 + H * kLdW1,
              kScrB1 = kScrW3 + NO * kLdW2, kScrB2 = kScrB1 + H, kScrEnd = kScrB2 + H;
static_assert(kScrEnd * 4 <= 3 * kRows * H * 2, "gradient scratch must fit in the H2 operand");
// PART 0: dW2, db2, dW3 (complete after the dW2 / dH1 stage: runs while the dW1 MMA is in flight); PART 1: dW1, db1.
template <int PART>
__device__ __forceinline__ void grad_out(uint8_t* sm0, const Smem& S, uint32_t tmem, const NetG& g, int obs_dim, int out_dim,
                                         float* __restrict__ grad, bool first) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cq = warp >> 2;
    float* scr = reinterpret_cast<float*>(sm0 + S.H2.base);       // H2 (dZ2) is dead after the dW2 / dH1 stage
    const uint32_t t0 = tmem + ((32u * q) << 16);
    const int o = 16 * q + lane;
    float v[8];
    if (PART == 0) {
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 32) {                  // dW2 [o][i]
            umma::tmem_ld8(t0 + cDW2 + c0 + 8 * cq, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) scr[kScrW2 + o * kLdW2 + c0 + 8 * cq + j] = v[j];
            }
        }
        if (cq < NO / 8) {                                     // dW3^T [k][a] -> [a][k]
            umma::tmem_ld8(t0 + cDW3 + 8 * cq, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) scr[kScrW3 + (8 * cq + j) * kLdW2 + o] = v[j];
            }
        } else if (cq == 2) {                                  // db2: the ones column of [dW2 | db2]
            umma::tmem_ld8(t0 + cDW2 + (uint32_t)H, v);
            if (lane < 16) scr[kScrB2 + o] = v[0];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < H * H / kThreads; ++u) {
            const int e = tid + u * kThreads;
            out_acc(grad + g.w2 + e, scr[kScrW2 + (e >> 6) * kLdW2 + (e & (H - 1))], first);
        }
        if (warp < out_dim) {
            out_acc(grad + g.w3 + warp * H + lane, scr[kScrW3 + warp * kLdW2 + lane], first);
            out_acc(grad + g.w3 + warp * H + 32 + lane, scr[kScrW3 + warp * kLdW2 + 32 + lane], first);
        }
        if (tid < H) out_acc(grad + g.b2 + tid, scr[kScrB2 + tid], first);
    } else {
        if (8 * cq < S.KXP) {                                  // dW1 [o][i]   (warp-uniform)
            umma::tmem_ld8(t0 + cDW1 + 8 * cq, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) scr[kScrW1 + o * kLdW1 + 8 * cq + j] = v[j];
            }
        }
        if (cq == 3) {                                         // db1: the ones column of [dW1 | db1]
            umma::tmem_ld8(t0 + cDW1 + (uint32_t)S.KXP, v);
            if (lane < 16) scr[kScrB1 + o] = v[0];
        }
        __syncthreads();
        if (lane < obs_dim) {
#pragma unroll
            for (int r = warp; r < H; r += kThreads / 32)
                out_acc(grad + g.w1 + (int64_t)r * obs_dim + lane, scr[kScrW1 + r * kLdW1 + lane], first);
        }
        if (tid < H) out_acc(grad + g.b1 + tid, scr[kScrB1 + tid], first);
    }
}

// Sum 32 per-lane values over the warp with 31 shuffles; lane j returns the total of v[j].
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1) {
        const bool upper = (lane & n) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = upper ? v[i] : v[i + n];
            const float keep = upper ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}

// forward of one trunk: X -> H1 -> H2 -> head accumulator D3 (TMEM); h1 / h2 of the thread's
// (row, 16 columns) stay in registers
__device__ __forceinline__ void trunk_forward(uint8_t* sm, uint8_t* sm0, const Smem& S, uint32_t tmem, Pipe& pipe,
                                              float (&h1)[kCols], float (&h2)[kCols]) {
    pipe.run([&] { gemm_kx(tmem + cD1, 128, H, S.X, 0, S.W1, 0, S.KXP); });
    epi_tanh(sm0, S.H1, tmem, cD1, reinterpret_cast<const float*>(sm + S.b1), h1);
    pipe.run([&] { gemm_ts(tmem, cD2, H, S.W2, 0); });                                 // A = H1 from TMEM
    epi_tanh(sm0, S.H2, tmem, cD2, reinterpret_cast<const float*>(sm + S.b2), h2);
    pipe.run([&] { gemm_ts(tmem, cD3, NO, S.W3, 0); });                                // A = H2 from TMEM
}

// backward of one trunk given dOut (S.DO / dof); writes all weight and bias gradients of the net.
// `weights_dead()` is called (all threads) once nothing reads the network's weight block any more.
template <class F>
__device__ __forceinline__ void trunk_backward(uint8_t* sm, uint8_t* sm0, const Smem& S, uint32_t tmem, Pipe& pipe,
                                               const NetG& g, int obs_dim, int out_dim, float* __restrict__ grad,
                                               const float (&h1)[kCols], const float (&h2)[kCols], bool first,
                                               F&& weights_dead) {
    pipe.issue([&] { gemm<kRows / 16, kWgradFull>(tmem + cDW3, 64, NO, S.H2, 1, S.DO, 1); });    // dW3^T = H2^T dOut
    float dz2[kCols];
    head_input_grad_compute(sm, S, out_dim, h2, dz2);          // overlaps the MMA (reads dof / w3f / registers only)
    pipe.wait();
    tstamp(16);
    head_input_grad_store(sm0, S, tmem, dz2);                  // H2 := dZ2 (the MMA no longer reads H2), T := dZ2
    tstamp(17);
    pipe.run([&] {
        gemm<kRows / 16, kWgradFull>(tmem + cDW2, 64, H + 8, S.H2, 1, S.H1, 1);                   // [dW2 | db2] = dZ2^T [H1 | 1]
        gemm_ts(tmem, cDH1, H, S.W2, 1);                                              // dH1 = dZ2 W2, A = dZ2 from TMEM
    });
    tstamp(18);
    weights_dead();
    epi_dtanh(sm0, S.H1, tmem, cDH1, h1);                                             // H1 := dZ1
    tstamp(19);
    pipe.issue([&] {
        gemm<kRows / 16, kWgradFull>(tmem + cDW1, 64, S.KXP + 8, S.H1, 1, S.X, 1);                // [dW1 | db1] = dZ1^T [X | 1]
    });
    grad_out<0>(sm0, S, tmem, g, obs_dim, out_dim, grad, first);     // dW2 / db2 / dW3 leave while the MMA runs
    pipe.wait();
    tstamp(20);
    grad_out<1>(sm0, S, tmem, g, obs_dim, out_dim, grad, first);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// State of the in-kernel grid barriers of the fused epoch kernel (self-resetting; launches are
// stream-ordered and every CTA of the grid is resident: cooperative launch, 1 CTA / SM).
// They live in the 128-byte control block in front of the caller's weight image (one per model instance, so
// two models updating on different streams never share barrier state); without an image: this global block.
struct EpochCtl {
    unsigned int arrive, depart;
    double ss[2];
};
constexpr size_t kCtlBytes = 128;
__device__ EpochCtl g_ep_ctl = {0u, 0u, {0.0, 0.0}};

struct AdamArgs {     // optimiser half of the fused single-GPU path
    float* params_w;
    float* grad_scratch;   // n_params + TS_PPO_GRAD_EXTRA folded values (loss sums are read back from here)
    float* exp_avg;
    float* exp_avg_sq;
    int64_t* step_count;
    float* stats;          // one row of TS_PPO_STATS_STRIDE floats per minibatch (nullable)
};

struct GridBarrier {   // monotonic counter: the k-th use waits for k * gridDim.x arrivals
    unsigned int target;
    unsigned int* ctr;
    // Split phase.  arrive(): release at gpu scope (cumulative over the bar.sync) -- a release waits for the
    // calling thread's OUTSTANDING LOADS too, so prefetches that should fly across the barrier are issued
    // between arrive() and wait().
    __device__ __forceinline__ void arrive() {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        }
    }
    __device__ __forceinline__ void wait() {
        if (threadIdx.x == 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            } while (seen < target);
        }
        __syncthreads();
    }
    __device__ __forceinline__ void sync() { arrive(); wait(); }
};

// Cross-GPU sum of one gradient element (fused all-reduce of the epoch kernel): the peers' (value, seq) packets of this
// step are polled CONCURRENTLY -- one load per missing peer and round, all in flight together -- so the wait is one NVLink
// latency, not world - 1 dependent ones; the sum runs in rank order (bit-identical on every rank).  Kept out of line: its
// register arrays must not weigh on the single-GPU path.
__device__ __noinline__ float peer_gather_sum(const unsigned long long* src0, int world, int rank, float g, unsigned int seq,
                                              size_t peer_stride) {
    float vals[tsb::kMaxPeers];
    unsigned int missing = ((1u << world) - 1u) & ~(1u << rank);
    unsigned long long t_start = 0ull;
    unsigned int spins = 0u;
    while (missing) {
        unsigned long long got[tsb::kMaxPeers];
#pragma unroll
        for (int r = 0; r < tsb::kMaxPeers; ++r) {
            if (missing & (1u << r))
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got[r]) : "l"(src0 + (size_t)r * peer_stride) : "memory");
        }
#pragma unroll
        for (int r = 0; r < tsb::kMaxPeers; ++r) {
            if ((missing & (1u << r)) && (unsigned int)(got[r] >> 32) == seq) {
                vals[r] = __uint_as_float((unsigned int)got[r]);
                missing &= ~(1u << r);
            }
        }
        if (missing && (++spins & 0xffffu) == 0u) {      // a peer that never shows up must not hang the GPU
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t_start == 0ull) t_start = now;
            else if (now - t_start > 20000000000ull) __trap();
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int r = 0; r < tsb::kMaxPeers; ++r) {
        if (r < world) sum += (r == rank) ? g : vals[r];
    }
    return sum;
}

struct TileIn { float xv[8]; float av[kMaxAct]; float rv[4]; };   // one thread's share of a tile's gathers

// The minibatches of one launch are [lo0 + m * mb_size, lo0 + (m + 1) * mb_size) for m < n_mb - 1 and
// [lo0 + (n_mb - 1) * mb_size, end) for the last one (Batch.split with merge_last, batch.py:1196-1215).
//
// EPOCH = false: n_mb = 1; write this CTA's partial gradient row and stop (multi-GPU: fold + all-reduce
// + ts_clip_adam_step follow).
// EPOCH = true: persistent over the n_mb optimiser steps of one pass over the rollout.  Per step: tile
// forward/backward -> grid barrier -> every CTA folds its slice of the gradient over the partial rows ->
// barrier (global sum of squares) -> clip + Adam on the slice -> barrier (parameters visible).  The gathers
// of the NEXT minibatch's tile are issued before the first barrier and land while the CTA waits.
template <bool EPOCH>
__global__ void __launch_bounds__(kThreads, 1) ppo_tc_kernel(
    const float* params, const ts_actor_critic_desc d, const ts_ppo_hparams hp,
    const float* __restrict__ obs, const float* __restrict__ act, const float* __restrict__ adv,
    const float* __restrict__ ret, const float* __restrict__ logp_old, const float* __restrict__ v_s,
    const int32_t* __restrict__ perm, int64_t lo0, int64_t mb_size, int64_t end, int n_mb, int64_t global_rows,
    const float* __restrict__ adv_moments, float* __restrict__ partials, const AdamArgs opt, uint8_t* wimg /* nullable */,
    const tsb::PeerArgs px) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint64_t s_wbar;
    __shared__ float s_coef, s_norm, s_step_size, s_bc2_sqrt;
    __shared__ int32_t s_row[kRows];
    __shared__ float s_part[4][128];      // actor loss epilogue: log-prob partials of the 4 action groups; optimiser half: fold partials
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = (int)gridDim.x;
    const int64_t width = d.n_params + TS_PPO_GRAD_EXTRA;
    // this CTA's private partial-gradient row: no cross-CTA atomics
    float* __restrict__ grad = partials + (size_t)blockIdx.x * (size_t)width;
    auto mb_lo = [&](int m) { return lo0 + (int64_t)m * mb_size; };
    auto mb_hi = [&](int m) { return m == n_mb - 1 ? end : lo0 + (int64_t)(m + 1) * mb_size; };
    auto mb_tiles = [&](int m) { return (mb_hi(m) - mb_lo(m) + kRows - 1) / kRows; };
    auto prefetch_rows = [&](int m, int64_t t) {   // dataset row of every tile row, ahead of its use
        if (tid < kRows) {
            const int64_t pos = mb_lo(m) + t * kRows + tid;
            s_row[tid] = pos < mb_hi(m) ? (perm ? __ldg(perm + pos) : (int32_t)pos) : 0;
        }
    };
    if ((int64_t)blockIdx.x < mb_tiles(0)) prefetch_rows(0, blockIdx.x);
    const uint32_t sbase = umma::smem_u32(sm);
    uint8_t* sm0 = sm - sbase;     // so that (sm0 + shared_address) is the generic pointer
    const Smem S = make_smem(d.obs_dim, sbase);
    const int A = d.act_dim;
    const NetG ga{d.a_w1, d.a_b1, d.a_w2, d.a_b2, d.a_w3, d.a_b3, d.a_logstd};
    const NetG gc{d.c_w1, d.c_b1, d.c_w2, d.c_b2, d.c_w3, d.c_b3, -1};
    const int64_t step0 = EPOCH ? *opt.step_count : 0;
    const unsigned int seq0 = (EPOCH && px.world > 1) ? *((volatile unsigned int*)px.hdr) : 0u;

    if (warp == 0) umma::tmem_alloc(&s_tmem, kTmemCols);
    if (tid == 0) { umma::mbar_init(&s_bar, 1); umma::mbar_init(&s_wbar, 1); umma::fence_mbar_init(); }
    for (int e = tid; e < 2 * 3 * kRows; e += kThreads) {   // the ones chunks of X and H1 (bf16 1.0 = 0x3F80 in piece 0)
        const int r = e % kRows, pc = (e / kRows) % 3;
        const Mat& M = e < 3 * kRows ? S.X : S.H1;
        const uint32_t c = e < 3 * kRows ? (uint32_t)S.KXP : (uint32_t)H;
        const uint32_t w = pc == 0 ? 0x3F803F80u : 0u;
        *reinterpret_cast<uint4*>(sm0 + (M.base + pc * M.part + moff((uint32_t)r, c, M.RS))) = make_uint4(w, w, w, w);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    Pipe pipe{&s_bar, 0u};
    EpochCtl* ctl = wimg != nullptr ? reinterpret_cast<EpochCtl*>(wimg - kCtlBytes) : &g_ep_ctl;
    GridBarrier gbar{0u, &ctl->arrive};
    // Weight staging.  With a weight image: one bulk copy (TMA engine) per network, issued as early as the
    // block is dead, completion on s_wbar.  Without: gather + split in the CTA (stage_weights).
    uint32_t wphase = 0u;
    bool critic_issued = false;
    auto issue_weights = [&](int net) {   // every earlier access to the weight block is ordered before this call
        if (wimg != nullptr && tid == 0) {
            umma::fence_proxy_async_all();
            umma::mbar_expect_tx(&s_wbar, S.wblk_bytes);
            umma::bulk_g2s(sbase + S.wblk, wimg + (size_t)net * S.wblk_bytes, S.wblk_bytes, &s_wbar);
        }
    };
    auto wait_weights = [&](const NetG& g, int out_dim) {
        if (wimg == nullptr) { stage_weights(sm, sm0, S, params, g, d.obs_dim, out_dim); return; }
        umma::mbar_wait(&s_wbar, wphase);
        wphase ^= 1u;
        if (g.ls >= 0 && tid < kMaxAct) {     // per-dimension constants of the diagonal Gaussian
            float* gs = reinterpret_cast<float*>(sm + S.red) + 64;
            const float sigma = expf(reinterpret_cast<const float*>(sm + S.ls)[tid]);
            gs[tid] = 1.0f / (sigma * sigma);
            gs[16 + tid] = logf(sigma) + 0.9189385332046727f;
        }
    };
    float* rowv = reinterpret_cast<float*>(sm + S.rowv);
    float* red = reinterpret_cast<float*>(sm + S.red);
    float* actt = reinterpret_cast<float*>(sm + S.act);

    // every gather of a tile in flight together (rows from s_row) ...
    auto load_inputs = [&](int nrows, TileIn& in) {
        chunk_load(kRows, d.obs_dim, S.KXP, [&](int r) {
            return r < nrows ? obs + (int64_t)s_row[r] * d.obs_dim : (const float*)nullptr;
        }, in.xv);
        if (tid < kRows) {
            const int64_t row = tid < nrows ? (int64_t)s_row[tid] : -1;
#pragma unroll
            for (int a = 0; a < kMaxAct; ++a) in.av[a] = (row >= 0 && a < A) ? __ldg(act + row * A + a) : 0.0f;
            in.rv[0] = in.rv[1] = in.rv[2] = in.rv[3] = 0.0f;
            if (row >= 0) { in.rv[0] = __ldg(adv + row); in.rv[1] = __ldg(ret + row); in.rv[2] = __ldg(logp_old + row); in.rv[3] = __ldg(v_s + row); }
        }
    };
    // ... and their conversion into the tile's operands
    auto store_inputs = [&](const TileIn& in) {
        chunk_store(sm0, S.X, kRows, S.KXP, in.xv);
        if (tid < kRows) {
#pragma unroll
            for (int a = 0; a < kMaxAct; ++a) actt[a * kRows + tid] = in.av[a];      // [a][r]: rows on consecutive banks
            rowv[tid] = in.rv[0]; rowv[kRows + tid] = in.rv[1]; rowv[2 * kRows + tid] = in.rv[2]; rowv[3 * kRows + tid] = in.rv[3];
        }
    };

    double beta1_pow = 1.0, beta2_pow = 1.0;
    if (EPOCH && tid == 256) { beta1_pow = pow(hp.beta1, (double)step0); beta2_pow = pow(hp.beta2, (double)step0); }
    bool staged = false;        // the tile's inputs were already stored by the previous step's prefetch
    tstamp(0);
    for (int m = 0; m < n_mb; ++m) {
        const int64_t lo = mb_lo(m), hi = mb_hi(m);
        const int64_t tiles = (hi - lo + kRows - 1) / kRows;
        // the loss is the mean over the GLOBAL minibatch: every rank contributes hi - lo rows of its own shard
        const ppo::Scalars sc = ppo::make_scalars(hp, EPOCH ? (hi - lo) * px.world : global_rows, adv_moments ? adv_moments + 2 * m : nullptr);
#ifdef TS_B200_DIAGNOSTICS
        if (tid == 0 && g_tc_timeline_on && (int)blockIdx.x == g_tc_timeline_on - 1) g_tc_timeline_gate = (n_mb == 1 || m == n_mb - 2);   // a step WITH barrier 3
#endif
        tstamp(22);
        if (EPOCH && tid == 256) {    // Adam bias corrections of this step, off the critical path
            beta1_pow *= hp.beta1; beta2_pow *= hp.beta2;          // beta^(step0 + m + 1)
            s_step_size = (float)(hp.lr / (1.0 - beta1_pow));
            s_bc2_sqrt = (float)sqrt(1.0 - beta2_pow);
        }
        bool next_rows_ready = false;   // s_row holds the rows of this CTA's first tile of minibatch m + 1
        for (int64_t t = blockIdx.x; t < tiles; t += P) {
            const bool first = (t == (int64_t)blockIdx.x);     // first tile of this CTA: gradients are stored, not added
            const int nrows = (int)tsb::imin((int64_t)kRows, hi - (lo + t * kRows));
            if (!staged) {
                TileIn in;
                load_inputs(nrows, in);
                store_inputs(in);
            }
            staged = false;

            // dataset rows of this CTA's NEXT tile: the (DRAM-latency) load of the permutation entry is issued here and
            // committed to s_row after the critic pass -- s_row itself was consumed by load_inputs above / one step ago
            int32_t next_row = 0;
            int next_kind = 0;                                   // 1: next tile of this minibatch, 2: first tile of the next one
            if (t + P < tiles) next_kind = 1;
            else if (m + 1 < n_mb && (int64_t)blockIdx.x < mb_tiles(m + 1)) next_kind = 2;
            if (next_kind != 0 && tid < kRows) {
                const int mm = next_kind == 1 ? m : m + 1;
                const int64_t pos = mb_lo(mm) + (next_kind == 1 ? t + P : (int64_t)blockIdx.x) * kRows + tid;
                next_row = pos < mb_hi(mm) ? (perm ? __ldg(perm + pos) : (int32_t)pos) : 0;
            }

            // ================= critic ================================================================
            float h1[kCols], h2[kCols];
            tstamp(1);
            if (!critic_issued) issue_weights(0);
            critic_issued = false;
            wait_weights(gc, 1);
            tstamp(2);
            trunk_forward(sm, sm0, S, tmem, pipe, h1, h2);
            tstamp(3);
            float vf_row = 0.0f;
            if (tid < kRows) {
                float v16[16], dv[kMaxAct];
                umma::tmem_ld16(tmem + ((32u * warp) << 16) + cD3, v16);
#pragma unroll
                for (int a = 0; a < kMaxAct; ++a) dv[a] = 0.0f;
                if (tid < nrows) {
                    const float value = v16[0] + reinterpret_cast<const float*>(sm + S.b3)[0];
                    ppo::critic_row(sc, value, rowv[kRows + tid], rowv[3 * kRows + tid], vf_row, dv[0]);
                }
                write_dout_row(sm, sm0, S, tid, dv);
                const float sdv = warp_sum(dv[0]);
                if (lane == 0) red[warp] = sdv;                    // db3 (critic), one slot per warp
            }
            tstamp(4);
            trunk_backward(sm, sm0, S, tmem, pipe, gc, d.obs_dim, 1, grad, h1, h2, first, [&] { issue_weights(1); });
            tstamp(5);
            __syncthreads();
            if (tid == 0) out_acc(grad + gc.b3, (red[0] + red[1]) + (red[2] + red[3]), first);
            // row ids of this CTA's next tile (loaded at the top of the tile; s_row was consumed by load_inputs long ago)
            if (next_kind != 0 && tid < kRows) s_row[tid] = next_row;
            if (next_kind == 2) next_rows_ready = true;

            // ================= actor =================================================================
            wait_weights(ga, A);
            tstamp(6);
            trunk_forward(sm, sm0, S, tmem, pipe, h1, h2);
            tstamp(7);
            // Actor loss epilogue on all 16 warps: thread (row r = 32 q + lane, group cq) owns the actions a = cq + 4 u -- the
            // four groups' log-prob partials meet in shared memory, every thread then evaluates the row's surrogate and writes
            // dOut / column sums of ITS actions only (the 128-thread version was a 2 us dependent chain on one warp per scheduler).
            float clip_row = 0.0f;
            {
                const int q = warp & 3, cq = warp >> 2;
                const int r = 32 * q + lane;
                const float* b3 = reinterpret_cast<const float*>(sm + S.b3);
                const float* inv_var = red + 64;      // 1 / sigma^2
                const float* logc = red + 80;         // log sigma + log sqrt(2 pi)
                float diff[4], d2v[4];
                float lpp = 0.0f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int a = cq + 4 * u;
                    diff[u] = 0.0f; d2v[u] = 0.0f;
                    if (a < A) {     // warp-uniform
                        const float mu = umma::tmem_ld1(tmem + ((32u * q) << 16) + cD3 + (uint32_t)a) + b3[a];
                        diff[u] = actt[a * kRows + r] - mu;
                        d2v[u] = diff[u] * diff[u] * inv_var[a];
                        lpp += fmaf(-0.5f, d2v[u], -logc[a]);      // log N(x; mu, sigma)
                    }
                }
                s_part[cq][r] = lpp;
                __syncthreads();
                const float lp = (s_part[0][r] + s_part[1][r]) + (s_part[2][r] + s_part[3][r]);
                float gl = 0.0f, obj = 0.0f;
                if (r < nrows) ppo::actor_row(sc, lp, rowv[2 * kRows + r], rowv[r], obj, gl);
                if (cq == 0) clip_row = obj;          // one group carries the row's objective into the loss sum
                float* dof = reinterpret_cast<float*>(sm + S.dof);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int a = cq + 4 * u;
                    if (a < A) {     // warp-uniform
                        const float g_mu = gl * diff[u] * inv_var[a];          // d loss / d mu_a
                        const float g_ls = gl * (d2v[u] - 1.0f);               // d loss / d logstd_a
                        dof[a * kRows + r] = g_mu;
                        uint32_t w0, w1, w2;
                        split3_pair(g_mu, 0.0f, w0, w1, w2);
                        uint8_t* pdo = sm0 + (S.DO.base + moff((uint32_t)r, (uint32_t)a, S.DO.RS));   // columns >= A stay zero (critic pass)
                        *reinterpret_cast<uint16_t*>(pdo) = (uint16_t)w0;
                        *reinterpret_cast<uint16_t*>(pdo + S.DO.part) = (uint16_t)w1;
                        *reinterpret_cast<uint16_t*>(pdo + 2 * S.DO.part) = (uint16_t)w2;
                        // column sums over the 32 rows of this warp: red[128 + 32 q + a] <- db3[a], red[128 + 32 q + 16 + a] <- dlogstd[a]
                        const float s_mu = warp_sum(g_mu), s_ls = warp_sum(g_ls);
                        if (lane == 0) { red[128 + 32 * q + a] = s_mu; red[128 + 32 * q + 16 + a] = s_ls; }
                    }
                }
            }
            tstamp(8);
            trunk_backward(sm, sm0, S, tmem, pipe, ga, d.obs_dim, A, grad, h1, h2, first, [] {});
            tstamp(9);

            // ================= loss sums + small gradients ===========================================
            const float s_clip = warp_sum(tid < kRows ? clip_row : 0.0f);
            const float s_vf = warp_sum(tid < kRows ? vf_row : 0.0f);
            if (lane == 0 && tid < kRows) { red[4 + warp] = s_clip; red[8 + warp] = s_vf; }
            __syncthreads();
            if (tid < A) {
                const float* cw = red + 128 + tid;
                out_acc(grad + ga.b3 + tid, (cw[0] + cw[32]) + (cw[64] + cw[96]), first);
                out_acc(grad + ga.ls + tid, (cw[16] + cw[48]) + (cw[80] + cw[112]) - sc.ent_coef * sc.inv_b * (float)nrows, first);
            }
            if (tid == 0) {
                float ent = 0.0f;     // entropy of the diagonal Gaussian: sum_a (0.5 + log sqrt(2 pi) + log sigma_a)
                for (int a = 0; a < A; ++a) ent += 0.5f + red[80 + a];
                float* ex = grad + d.n_params;
                out_acc(ex + 0, (red[4] + red[5]) + (red[6] + red[7]), first);
                out_acc(ex + 1, (red[8] + red[9]) + (red[10] + red[11]), first);
                out_acc(ex + 2, ent * (float)nrows, first);
                out_acc(ex + 3, (float)nrows, first);
            }
            __syncthreads();
        }
        if (!EPOCH) break;

        // ---- optimiser half of the step --------------------------------------------------------------
        __shared__ double s_red[4];
        const int Pm = (int)tsb::imin((int64_t)P, tiles);           // partial rows written for this minibatch
        const int64_t slice = (width + P - 1) / P;
        const int64_t i0 = (int64_t)blockIdx.x * slice;
        const int64_t i1 = tsb::imin(i0 + slice, width);
        const int e = tid & 127, q = tid >> 7;
        const int64_t step = step0 + m + 1;
        double* ss_cur = &ctl->ss[m & 1];
        tstamp(10);
        // gathers of this CTA's tile of the next minibatch: in flight across the barrier
        TileIn pin;
        const bool pre = (m + 1 < n_mb) && (int64_t)blockIdx.x < mb_tiles(m + 1);
        if (pre && !next_rows_ready) prefetch_rows(m + 1, blockIdx.x);
        gbar.arrive();
        if (pre) {
            const int nrows_next = (int)tsb::imin((int64_t)kRows, mb_hi(m + 1) - (mb_lo(m + 1) + (int64_t)blockIdx.x * kRows));
            load_inputs(nrows_next, pin);
        }
        gbar.wait();                                                // every partial row is complete
        tstamp(11);
        if (pre) { store_inputs(pin); staged = true; }
        tstamp(23);
        // fold: thread (e, q) sums rows q, q + 4, ... of element i0 + e [+ 128, ...]; 32 loads in flight
        double ss = 0.0;
        for (int64_t c = i0; c < i1; c += 128) {
            const int64_t i = c + e;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (i < i1) {
                for (int p0 = q; p0 < Pm; p0 += 128) {
                    float v[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) {
                        const int p = p0 + 4 * u;
                        v[u] = p < Pm ? __ldcg(partials + (int64_t)p * width + i) : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < 32; u += 4) { acc[0] += v[u]; acc[1] += v[u + 1]; acc[2] += v[u + 2]; acc[3] += v[u + 3]; }
                }
            }
            s_part[q][e] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
            __syncthreads();
            tstamp(24);
            if (q == 0 && i < i1) {
                float g = (s_part[0][e] + s_part[1][e]) + (s_part[2][e] + s_part[3][e]);
                if (px.world > 1) {
                    // ---- cross-GPU sum of this element over NVLink peer memory (fused all-reduce) -----------
                    // push (value, seq) into every peer's receive slot of this rank, then gather the peers'
                    // packets from the local buffer and add in rank order (bit-identical on every rank)
                    const unsigned int seq = seq0 + (unsigned int)m + 1u;
                    const size_t slot = (size_t)(seq & 1u) * (size_t)width + (size_t)i;
                    const unsigned long long pkt = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(g);
                    for (int r = 0; r < px.world; ++r) {
                        if (r == px.rank) continue;
                        unsigned long long* dst = px.recv[r] + (size_t)px.rank * 2u * (size_t)width + slot;
                        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(pkt) : "memory");
                    }
                    g = peer_gather_sum(px.recv[px.rank] + slot, px.world, px.rank, g, seq, 2u * (size_t)width);
                }
                opt.grad_scratch[i] = g;
                if (i < d.n_params) ss += (double)g * (double)g;
            }
            __syncthreads();
        }
        if (tid < 128) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss += tsb::shfl_xor_f64(ss, off);
            if (lane == 0) s_red[warp] = ss;
        }
        __syncthreads();
        tstamp(12);
        if (tid == 0) atomicAdd(ss_cur, (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]));
        // the optimiser state of this thread's elements: loads in flight across the barrier
        const bool one_pass = slice <= kThreads;      // (always, unless the network is far larger than the grid)
        const int64_t i_own = i0 + tid;
        const bool own = one_pass && i_own < i1 && i_own < d.n_params;
        float pv_own = 0.f, m_own = 0.f, v_own = 0.f;
        gbar.arrive();
        if (own) { pv_own = opt.params_w[i_own]; m_own = opt.exp_avg[i_own]; v_own = opt.exp_avg_sq[i_own]; }
        gbar.wait();                                                // global sum of squares is complete
        tstamp(13);
        if (tid == 0) {
            const float total_norm = (float)sqrt(*((volatile double*)ss_cur));
            float coef = 1.0f;
            if (hp.max_grad_norm > 0.0) {   // torch.nn.utils.clip_grad_norm_
                coef = (float)hp.max_grad_norm / (total_norm + 1e-6f);
                coef = fminf(coef, 1.0f);
            }
            s_coef = coef; s_norm = total_norm;
            if (blockIdx.x == 0) ctl->ss[(m + 1) & 1] = 0.0;       // next step's accumulator (idle until barrier 3)
        }
        __syncthreads();
        tstamp(25);
        const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
        const float w1 = (float)(1.0 - hp.beta1), w2 = (float)(1.0 - hp.beta2);
        const float beta2 = (float)hp.beta2, adam_eps = (float)hp.adam_eps, wd = (float)hp.weight_decay;
        auto adam_elem = [&](int64_t i, float g, float pv, float mm, float v) {
            g *= coef;
            if (wd != 0.0f) g = fmaf(wd, pv, g);
            mm = mm + w1 * (g - mm);                    // exp_avg.lerp_(grad, 1 - beta1)
            v = v * beta2 + w2 * g * g;                 // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(v) / bc2_sqrt + adam_eps;
            pv = pv - step_size * (mm / denom);         // addcdiv_(exp_avg, denom, -step_size)
            opt.exp_avg[i] = mm; opt.exp_avg_sq[i] = v; opt.params_w[i] = pv;
            if (wimg != nullptr) img_scatter(d, S, sbase, i, pv, wimg);
        };
        if (one_pass) {
            if (own) adam_elem(i_own, opt.grad_scratch[i_own], pv_own, m_own, v_own);
        } else {
            for (int64_t i = i0 + tid; i < i1 && i < d.n_params; i += kThreads)
                adam_elem(i, opt.grad_scratch[i], opt.params_w[i], opt.exp_avg[i], opt.exp_avg_sq[i]);
        }
        tstamp(26);
        if (wimg != nullptr) umma::fence_proxy_async_all();      // image stores (generic proxy) before the peers' bulk copies
        tstamp(14);
        const bool more = m + 1 < n_mb;
        if (more) gbar.arrive();                                    // barrier 3 (updated parameters visible to every CTA) ...
        if (tid == 32 && blockIdx.x == 0) {                         // ... the loss table row is written under it, off tid 0's poll
            // the folded loss sums were written by the owner of the last slice before barrier 2; nothing rewrites them before the
            // next step's fold, i.e. after every CTA passed the NEXT barrier 1
            const float* ex = opt.grad_scratch + d.n_params;
            const float e0 = __ldcg(ex), e1 = __ldcg(ex + 1), e2 = __ldcg(ex + 2), e3 = __ldcg(ex + 3);
            const float rows = e3 > 0.0f ? e3 : 1.0f;
            const float clip_loss = -e0 / rows, vf_loss = e1 / rows, ent_loss = e2 / rows;
            if (opt.stats) {
                float* sr = opt.stats + (int64_t)m * TS_PPO_STATS_STRIDE;
                sr[0] = clip_loss + (float)hp.vf_coef * vf_loss - (float)hp.ent_coef * ent_loss;
                sr[1] = clip_loss; sr[2] = vf_loss; sr[3] = ent_loss;
                sr[4] = s_norm; sr[5] = e3; sr[6] = 0.0f; sr[7] = 0.0f;
            }
        }
        if (more) {
            gbar.wait();
            if (pre) { issue_weights(0); critic_issued = true; }
        }
        tstamp(15);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
    if (EPOCH && tid == 0) {
        if (blockIdx.x == 0) {
            *opt.step_count = step0 + n_mb;
            if (px.world > 1) *px.hdr = seq0 + (unsigned int)n_mb;
        }
        __threadfence();
        if (atomicAdd(&ctl->depart, 1u) == gridDim.x - 1) {
            ctl->arrive = 0u; ctl->depart = 0u; ctl->ss[0] = 0.0; ctl->ss[1] = 0.0;
            __threadfence();
        }
    }
}


// ---- forward-only kernels (value pass / log-prob pass), persistent over 128-row tiles -----------
// TS mode: the activations (X, H1) are the A operand IN TENSOR MEMORY -- the staging / epilogue threads split
// them to bf16x3 and tcgen05.st them to their own TMEM lane (lane = row), so an MMA reads only the 2 KB weight
// operand from shared memory and runs at its 32-cycle math floor instead of being paced by 6 KB of operand reads
// (ncu, SS mode: tensor pipe busy 76 % of the pass).  Two tiles are in flight per CTA (slots 0 / 1: own
// accumulator and operand columns): the MMAs of one slot run while the 16 warps do the other slot's epilogues, and
// the next tiles' rows are loaded from global memory one stage ahead of their conversion.
// TMEM columns of a slot: D (64: layer-1, then layer-2 accumulator) | T (96: X pieces, then H1 pieces).
struct SmemF {
    int KXP;
    Mat W1, W2;
    uint32_t w3f, b1, b2, b3, ls, part;
    uint32_t total;
};
__host__ __device__ inline SmemF make_smem_f(int obs_dim, uint32_t sbase) {
    SmemF s;
    s.KXP = (obs_dim + 15) & ~15;
    uint32_t o = 0;
    auto mat = [&](Mat& m, int rows, int cols) {
        m.base = sbase + o; m.part = mat_bytes(rows, cols); m.RS = (uint32_t)(cols / 8) * 128u; o += 3u * m.part;
    };
    mat(s.W1, H, s.KXP);
    mat(s.W2, H, H);
    s.w3f = o;  o += kMaxAct * H * 4;
    s.b1 = o;   o += H * 4;
    s.b2 = o;   o += H * 4;
    s.b3 = o;   o += kMaxAct * 4;
    s.ls = o;   o += kMaxAct * 4;
    s.part = o; o += 4 * kRows * kMaxAct * 4;   // head partial sums of the four column groups
    s.total = o;
    return s;
}
constexpr uint32_t kSlotCols = 160, kSlotD = 0, kSlotT = 64;     // 2 slots -> 320 columns (512 allocated)

// all threads: publish smem / TMEM operands, retire TMEM reads, then warp 0 issues `f` and commits to `bar` (no wait)
template <class F>
__device__ __forceinline__ void mma_issue(uint64_t* bar, F&& f) {
    umma::tmem_wait_st();
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp_u == 0) {
        umma::fence_after_sync();
        f();
        if (umma::elect_one()) umma::mma_commit(bar);
        __syncwarp();
    }
}
__device__ __forceinline__ void mma_wait(uint64_t* bar, uint32_t& phase) {
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
}
// 16 consecutive values of the calling thread's row -> bf16x3 -> 8 packed columns per piece of its TMEM lane
__device__ __forceinline__ void store_row16_tmem(uint32_t t_lane, uint32_t col0, uint32_t part_cols, const float* v) {
    uint32_t w0[8], w1[8], w2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3_pair(v[2 * j], v[2 * j + 1], w0[j], w1[j], w2[j]);
    umma::tmem_st8(t_lane + col0, w0);
    umma::tmem_st8(t_lane + col0 + part_cols, w1);
    umma::tmem_st8(t_lane + col0 + 2u * part_cols, w2);
}

// MODE 0: out0[r] = critic(in0[r]) and (if in1) out1[r] = critic(in1[r])      (a2c.py:123-126)
// MODE 1: out0[r] = log N(in1[r] | mu(in0[r]), exp(logstd)), out1 = mu (nullable)   (ppo.py:157-161)
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) forward_tc_kernel(
    const float* __restrict__ params, const ts_actor_critic_desc d, const float* __restrict__ in0,
    float* __restrict__ out0, const float* __restrict__ in1, float* __restrict__ out1, int64_t n) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[2][2];      // [slot][layer]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cq = warp >> 2;
    const uint32_t sbase = umma::smem_u32(sm);
    uint8_t* sm0 = sm - sbase;
    const SmemF S = make_smem_f(d.obs_dim, sbase);
    const int out_dim = MODE == 0 ? 1 : d.act_dim;
    const NetG g = MODE == 0 ? NetG{d.c_w1, d.c_b1, d.c_w2, d.c_b2, d.c_w3, d.c_b3, -1}
                             : NetG{d.a_w1, d.a_b1, d.a_w2, d.a_b2, d.a_w3, d.a_b3, d.a_logstd};
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        umma::mbar_init(&s_bar[0][0], 1); umma::mbar_init(&s_bar[0][1], 1);
        umma::mbar_init(&s_bar[1][0], 1); umma::mbar_init(&s_bar[1][1], 1);
        umma::fence_mbar_init();
    }
    // weights: W1, W2 as tensor-core B operands (shared memory); head weights / biases as fp32
    stage_chunks(sm0, S.W1, H, d.obs_dim, S.KXP, [&](int o) { return params + g.w1 + (int64_t)o * d.obs_dim; });
    stage_chunks(sm0, S.W2, H, H, H, [&](int o) { return params + g.w2 + (int64_t)o * H; });
    float* w3f = reinterpret_cast<float*>(sm + S.w3f);
    float* b1 = reinterpret_cast<float*>(sm + S.b1);
    float* b2 = reinterpret_cast<float*>(sm + S.b2);
    float* b3 = reinterpret_cast<float*>(sm + S.b3);
    float* ls = reinterpret_cast<float*>(sm + S.ls);
    float* part = reinterpret_cast<float*>(sm + S.part);
    for (int e = tid; e < kMaxAct * H; e += kThreads) w3f[e] = (e >> 6) < out_dim ? __ldg(params + g.w3 + e) : 0.0f;
    for (int e = tid; e < H; e += kThreads) { b1[e] = __ldg(params + g.b1 + e); b2[e] = __ldg(params + g.b2 + e); }
    if (tid < kMaxAct) {
        b3[tid] = tid < out_dim ? __ldg(params + g.b3 + tid) : 0.0f;
        // sigma_a = exp(logstd_a), once per CTA; the log-prob keeps torch's expression order (normal_logp_term)
        ls[tid] = (MODE == 1 && tid < out_dim) ? expf(__ldg(params + g.ls + tid)) : 1.0f;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t t_lane = tmem + ((32u * q) << 16);       // this thread's TMEM lane group
    uint32_t ph[2][2] = {{0u, 0u}, {0u, 0u}};
    const uint32_t xcols = (uint32_t)S.KXP / 2u;            // packed columns of one X piece (8 or 16)

    const int64_t tiles_per = (n + kRows - 1) / kRows;
    const int64_t tiles = tiles_per * ((MODE == 0 && in1) ? 2 : 1);
    const int64_t my_n = tiles > (int64_t)blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA
    auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + k * gridDim.x; };
    auto tile_src = [&](int64_t t, const float*& src, int64_t& row0, int& nrows, bool& second) {
        second = t >= tiles_per;
        row0 = (second ? t - tiles_per : t) * kRows;
        nrows = (int)tsb::imin((int64_t)kRows, n - row0);
        src = (MODE == 0 && second) ? in1 : in0;
    };
    // X staging: thread (q, cq, lane) owns row 32 q + lane, columns [16 cq, 16 cq + 16) (warps with 16 cq >= KXP idle)
    const bool x_owner = 16 * cq < S.KXP;
    auto x_load = [&](int64_t k, float (&xv)[16]) {
        if (!x_owner) return;
        const float* src; int64_t row0; int nrows; bool second;
        tile_src(tile_of(k), src, row0, nrows, second);
        const int r = 32 * q + lane;
        const float* row = src + (row0 + r) * d.obs_dim;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = 16 * cq + j;
            xv[j] = (r < nrows && c < d.obs_dim) ? __ldg(row + c) : 0.0f;
        }
    };
    auto x_store = [&](int slot, const float (&xv)[16]) {
        if (x_owner) store_row16_tmem(t_lane, kSlotCols * slot + kSlotT + 8u * cq, xcols, xv);
    };
    auto l1 = [&](int slot) {
        mma_issue(&s_bar[slot][0], [&] {
            const uint32_t dcol = tmem + kSlotCols * slot + kSlotD, acol = tmem + kSlotCols * slot + kSlotT;
            const uint32_t idesc = umma::idesc_bf16(128, H, 0, 0);
            if (S.KXP == 16) umma::gemm_bf16x3_ts_warp<1>(dcol, acol, xcols, S.W1.base, S.W1.part, 128u, S.W1.RS, 256u, idesc);
            else umma::gemm_bf16x3_ts_warp<2>(dcol, acol, xcols, S.W1.base, S.W1.part, 128u, S.W1.RS, 256u, idesc);
        });
    };
    auto epi1_l2 = [&](int slot) {     // h1 = tanh(D + b1) -> T (3 x 32 packed columns); issue layer 2 into D
        mma_wait(&s_bar[slot][0], ph[slot][0]);
        {
            const uint32_t c0 = (uint32_t)kCols * cq;
            float h[kCols];
            umma::tmem_ld16(t_lane + kSlotCols * slot + kSlotD + c0, h);
#pragma unroll
            for (int j = 0; j < kCols; ++j) h[j] = tanh_mufu(h[j] + b1[c0 + j]);
            store_row16_tmem(t_lane, kSlotCols * slot + kSlotT + 8u * cq, (uint32_t)H / 2u, h);
        }
        mma_issue(&s_bar[slot][1], [&] {
            umma::gemm_bf16x3_ts_warp<H / 16>(tmem + kSlotCols * slot + kSlotD, tmem + kSlotCols * slot + kSlotT, (uint32_t)H / 2u,
                                              S.W2.base, S.W2.part, 128u, S.W2.RS, 256u, umma::idesc_bf16(128, H, 0, 0));
        });
    };
    auto epi2 = [&](int slot, int64_t k) {   // h2 = tanh(D + b2) in registers; head = h2 . W3^T (K = 64 split over the four column groups)
        const float* src; int64_t row0; int nrows; bool second;
        tile_src(tile_of(k), src, row0, nrows, second);
        mma_wait(&s_bar[slot][1], ph[slot][1]);
        {
            const uint32_t r = 32u * q + lane, c0 = (uint32_t)kCols * cq;
            float v[kCols];
            umma::tmem_ld16(t_lane + kSlotCols * slot + kSlotD + c0, v);
#pragma unroll
            for (int j = 0; j < kCols; ++j) v[j] = tanh_mufu(v[j] + b2[c0 + j]);
            for (int a = 0; a < out_dim; ++a) {
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < kCols; ++j) acc = fmaf(v[j], w3f[a * H + c0 + j], acc);
                part[(cq * kMaxAct + a) * kRows + r] = acc;     // [column group][a][row]: rows on consecutive banks
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        if (tid < nrows) {
            const int r = tid;
            if (MODE == 0) {
                (second ? out1 : out0)[row0 + r] = (part[r] + part[kMaxAct * kRows + r]) +
                                                   (part[2 * kMaxAct * kRows + r] + part[3 * kMaxAct * kRows + r]) + b3[0];
            } else {
                float lp = 0.0f;
                for (int a = 0; a < out_dim; ++a) {
                    const float* pa = part + a * kRows + r;
                    const float mu = (pa[0] + pa[kMaxAct * kRows]) + (pa[2 * kMaxAct * kRows] + pa[3 * kMaxAct * kRows]) + b3[a];
                    lp += ppo::normal_logp_term(__ldg(in1 + (row0 + r) * out_dim + a), mu, ls[a]);
                    if (out1) out1[(row0 + r) * out_dim + a] = mu;
                }
                out0[row0 + r] = lp;
            }
        }
        __syncthreads();     // `part` is free again; every thread has read this slot's D (layer 1 of the next tile may overwrite it)
    };

    float xa[16], xb[16];
    if (my_n > 0) { x_load(0, xa); x_store(0, xa); l1(0); }
    if (my_n > 1) { x_load(1, xb); x_store(1, xb); l1(1); }
    for (int64_t k = 0; k < my_n; k += 2) {
        const bool hasB = k + 1 < my_n, nextA = k + 2 < my_n, nextB = k + 3 < my_n;
        if (nextA) x_load(k + 2, xa);              // global loads fly under the epilogues below
        epi1_l2(0);
        if (nextB) x_load(k + 3, xb);
        if (hasB) epi1_l2(1);                      // layer 2 of slot 0 runs on the tensor core meanwhile
        epi2(0, k);
        // slot 0 is completely retired (its T columns were last read by layer 2, its D by the epilogue above)
        if (nextA) { x_store(0, xa); l1(0); }
        if (hasB) epi2(1, k + 1);
        if (nextB) { x_store(1, xb); l1(1); }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

#ifdef TS_B200_DIAGNOSTICS
extern "C" int ts_tc_timeline(int32_t enable, uint64_t* out32 /* host, nullable */) {
    int on = enable;
    TS_CUDA(cudaMemcpyToSymbol(g_tc_timeline_on, &on, sizeof(int)));
    if (out32) TS_CUDA(cudaMemcpyFromSymbol(out32, g_tc_timeline, 32 * sizeof(unsigned long long)));
    return 0;
}
#endif

namespace tsb {
bool tc_supported(const ts_actor_critic_desc& d) {
    // tanh trunks, Gaussian head, separate actor / critic parameters (a shared trunk = aliased offsets would break
    // the store-not-add gradient write-out); everything else runs the fp32 SIMT kernels
    return d.hidden == H && d.obs_dim >= 1 && d.obs_dim <= 32 && d.act_dim >= 1 && d.act_dim <= kMaxAct && d.flags == 0 &&
           d.a_w1 != d.c_w1 && d.a_logstd >= 0;
}

static int configure_ppo_smem(size_t smem) {
    static size_t configured[kMaxDevices] = {};       // the opt-in is a per-device function attribute
    const int dev = device_ordinal();
    if (smem > configured[dev]) {
        TS_CUDA(cudaFuncSetAttribute(ppo_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS_CUDA(cudaFuncSetAttribute(ppo_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    return 0;
}

int launch_ppo_grad_tc(const float* params, const ts_actor_critic_desc& d, const ts_ppo_hparams& hp, const float* obs,
                       const float* act, const float* adv, const float* ret, const float* logp_old, const float* v_s,
                       const int32_t* perm, int64_t lo, int64_t hi, int64_t global_rows, const float* adv_moments,
                       float* grad, cudaStream_t st) {
    const size_t smem = make_smem(d.obs_dim, 0).total;
    if (int e = configure_ppo_smem(smem)) return e;
    const int64_t tiles = (hi - lo + kRows - 1) / kRows;
    const unsigned grid = (unsigned)imin(tiles, num_sms());
    ppo_tc_kernel<false><<<grid, kThreads, smem, st>>>(params, d, hp, obs, act, adv, ret, logp_old, v_s, perm, lo, hi - lo, hi, 1,
                                                       global_rows, adv_moments, grad, AdamArgs{}, (uint8_t*)nullptr, PeerArgs{});
    return check_launch("ts_ppo_grad(tc)");
}

// n_mb consecutive optimiser steps (minibatch fwd/bwd + gradient fold + clip + Adam + stats each) in ONE
// cooperative launch: every CTA is resident (grid <= #SMs, 1 CTA / SM), so the in-kernel grid barriers are safe.
int launch_ppo_epoch_tc(float* params, const ts_actor_critic_desc& d, const ts_ppo_hparams& hp, const float* obs,
                        const float* act, const float* adv, const float* ret, const float* logp_old, const float* v_s,
                        const int32_t* perm, int64_t lo0, int64_t mb_size, int64_t end, int n_mb, const float* adv_moments,
                        float* partials, float* grad_scratch, float* exp_avg, float* exp_avg_sq, int64_t* step_count,
                        float* stats, void* weight_image, const PeerArgs& px, cudaStream_t st) {
    const size_t smem = make_smem(d.obs_dim, 0).total;
    if (int e = configure_ppo_smem(smem)) return e;
    const int64_t last = end - (lo0 + (int64_t)(n_mb - 1) * mb_size);
    const int64_t widest = n_mb > 1 ? (mb_size > last ? mb_size : last) : last;
    const unsigned grid = (unsigned)imin((widest + kRows - 1) / kRows, num_sms());
    const AdamArgs opt{params, grad_scratch, exp_avg, exp_avg_sq, step_count, stats};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int64_t zero = 0;
    // scratch layout: [128-byte control block (grid-barrier state, zero between launches)][critic block][actor block]
    uint8_t* wimg = weight_image ? static_cast<uint8_t*>(weight_image) + kCtlBytes : nullptr;
    if (wimg) {   // (re)build the pre-split image from the current parameters: the host may have changed them
        const Smem S = make_smem(d.obs_dim, 0);
        TS_CUDA(cudaMemsetAsync(wimg, 0, 2 * (size_t)S.wblk_bytes, st));
        weight_image_build_kernel<<<(unsigned)((d.n_params + 255) / 256), 256, 0, st>>>(params, d, wimg);
        if (int e = check_launch("ts_ppo_update(weight image)")) return e;
    }
    TS_CUDA(cudaLaunchKernelEx(&cfg, ppo_tc_kernel<true>, (const float*)params, d, hp, obs, act, adv, ret, logp_old, v_s, perm,
                               lo0, mb_size, end, n_mb, zero, adv_moments, partials, opt, wimg, px));
    return check_launch("ts_ppo_epoch(tc)");
}

int64_t weight_image_bytes(const ts_actor_critic_desc& d) {
    return tc_supported(d) ? (int64_t)kCtlBytes + 2 * (int64_t)make_smem(d.obs_dim, 0).wblk_bytes : 0;
}

int launch_forward_tc(int mode, const float* params, const ts_actor_critic_desc& d, const float* in0, float* out0,
                      const float* in1, float* out1, int64_t n, cudaStream_t st) {
    const size_t smem = make_smem_f(d.obs_dim, 0).total;
    static size_t configured[kMaxDevices][2] = {};
    const int dev = device_ordinal();
    if (smem > configured[dev][mode]) {
        if (mode == 0) TS_CUDA(cudaFuncSetAttribute(forward_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else TS_CUDA(cudaFuncSetAttribute(forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev][mode] = smem;
    }
    const int64_t tiles = ((n + kRows - 1) / kRows) * ((mode == 0 && in1) ? 2 : 1);
    const unsigned grid = (unsigned)imin(tiles, num_sms());
    if (mode == 0) forward_tc_kernel<0><<<grid, kThreads, smem, st>>>(params, d, in0, out0, in1, out1, n);
    else forward_tc_kernel<1><<<grid, kThreads, smem, st>>>(params, d, in0, out0, in1, out1, n);
    return check_launch(mode == 0 ? "ts_critic_forward(tc)" : "ts_actor_logp(tc)");
}
}  // namespace tsb
