// Shared declarations of the sm_100a kernels behind include/fo_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

typedef __nv_bfloat16 bf16;

namespace fo {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define FO_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -2;                                                                            \
        }                                                                                         \
    } while (0)
#define FO_CHECK(cond, ...)                                                                       \
    do {                                                                                          \
        if (!(cond)) { fo::set_error(__VA_ARGS__); return -1; }                                   \
    } while (0)
#define FO_TRY(expr) do { int _r = (expr); if (_r != 0) return _r; } while (0)

extern long long g_launches;      // kernels enqueued by this library (fo_stats.kernel_launches)
#define FO_LAUNCHED() (++fo::g_launches)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- dtype helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- implicit-GEMM row gather (conv2 of the subsampling, adapter conv) --------------------------
// Element (row, k) of the A operand lives at  A + (row_off(row) + seg_off(k / seg_len)) * seg_len + k % seg_len
// with row_off(row) = (row / (d1*d0)) * s2 + ((row / d0) % d1) * s1 + (row % d0) * s0
// and  seg_off(s)   = (s / seg_w) * seg_s1 + (s % seg_w) * seg_s0.
// plain row-major A is the special case seg_len = K, d0 = d1 = 1, s2 = 1.
struct AGather {
    int seg_len;          // contiguous run length along k (channels)
    int d0, d1;           // row index decomposition
    long long s0, s1, s2; // strides of the decomposition, in units of seg_len elements
    int seg_w;            // segments per kernel row (kw count); 0 => plain
    long long seg_s0, seg_s1;
};
inline AGather plain_rows(int K) { AGather g{K, 1, 1, 0, 0, 1, 0, 0, 0}; return g; }

__host__ __device__ inline long long gather_row_off(const AGather& g, int row) {
    return (long long)(row / (g.d1 * g.d0)) * g.s2 + (long long)((row / g.d0) % g.d1) * g.s1 +
           (long long)(row % g.d0) * g.s0;
}
__host__ __device__ inline long long gather_seg_off(const AGather& g, int seg) {
    return g.seg_w ? (long long)(seg / g.seg_w) * g.seg_s1 + (long long)(seg % g.seg_w) * g.seg_s0 : 0;
}

// ---- GEMM epilogue ---------------------------------------------------------------------------
// C[m][n] = act((acc + bias[n]) * scale) (+ residual[m][n]);  written as fp32 (c_f32) and/or as the
// activation type (c_act).  residual may alias c_f32 (in-place residual stream).
struct Epilogue {
    const float* bias = nullptr;
    const float* residual = nullptr;  // fp32, leading dimension ldc
    float* c_f32 = nullptr;
    void* c_act = nullptr;            // float* or bf16* depending on the context dtype
    int ldc = 0;
    int relu = 0;
    float scale = 1.0f;
};

// C[M,N] = A[M,K] * W[N,K]^T.  A/W are TIn (float or bf16), accumulate fp32.
template <typename TIn>
int gemm_simt(const TIn* A, const AGather& ga, const TIn* W, int M, int N, int K, const Epilogue& ep,
              cudaStream_t st);

// tcgen05 + TMA path (bf16 only).  Returns 1 if the shape is not supported (caller falls back
// to gemm_simt), <0 on error.
int gemm_tc_init();
int gemm_tc(const bf16* A, const AGather& ga, const bf16* W, int M, int N, int K, const Epilogue& ep,
            int split_k, cudaStream_t st);

// ---- frontend --------------------------------------------------------------------------------
struct FbankParams {
    int frame_len, frame_shift, fft_size, n_mel;
    const float* window;   // frame_len
    const float* mel;      // n_mel x (fft_size/2+1), zero outside each triangle
    const int* mel_lo;     // first non-zero fft bin per mel filter
    const int* mel_hi;     // one past the last non-zero bin
};
// Streaming: per session s = ids[b]: samples = [carry[s] (frame_len-frame_shift) | new chunk], m = frames_per_chunk
// frames; ring[s] <- [last ctx frames | new frames]; optional copy of the ring block to feats_out.
int fbank_stream(const FbankParams& p, const int32_t* ids, int n, const void* pcm, int pcm_is_i16, float scale,
                 int frames_per_chunk, int ctx_frames, float* carry, float* ring, float* feats_out,
                 cudaStream_t st);
int fbank_offline(const FbankParams& p, const void* pcm, int pcm_is_i16, int B, long long n_samples, float scale,
                  float* out, cudaStream_t st);

// ---- elementwise / normalisation ----------------------------------------------------------------
// CMVN + Conv2d(1->C,3,2) + ReLU, channels-last output c1[b][t1][f1][c]
template <typename TA>
int cmvn_conv1(const float* feats, int B, int T, int F, const float* mean, const float* istd, const float* w1,
               const float* b1, int C, TA* c1, cudaStream_t st);
// LayerNorm over the last dim of x (M, D) fp32.  y_act (activation type) and/or y_f32 outputs.
// act: 0 none, 1 relu, 2 gelu(erf); result multiplied by out_scale after the activation.
template <typename TA>
int layer_norm(const float* x, int M, int D, const float* gamma, const float* beta, float eps, int act,
               float out_scale, TA* y_act, float* y_f32, cudaStream_t st);
// scale-copy (input-layer "none"): y = x * s
int scale_rows(const float* x, float* y, long long n, float s, cudaStream_t st);
// adapter staging: xin[b][0..k-2] = cache (or 0), xin[b][k-1+i] = enc_out[b][i] (zeroed where mask==0);
// new_cache = last k-1 rows of xin (fp32).  cache layout (slot, k-1, D) time-major.
template <typename TA>
int adapter_stage(const float* enc_out, const uint8_t* mask, int B, int T, int D, int km1,
                  const int32_t* ids, float* slot_cache, int32_t* slot_valid,       // slot-resident (ids != null)
                  const float* cache_in, float* cache_out,                          // explicit (B, D, km1) layout
                  TA* xin, cudaStream_t st);
int subsample_mask(const int32_t* ilens, int B, int T, int T2, uint8_t* mask2, int32_t* ilens2, cudaStream_t st);
int stride2_mask(const uint8_t* mask, int B, int T, int To, uint8_t* out, cudaStream_t st);

// ---- attention -----------------------------------------------------------------------------------
struct AttnStream {
    const int32_t* ids;       // (n) session slots
    const int32_t* n_frames;  // per slot: encoder frames appended before this step
    const int32_t* pe_index;  // per slot
    int n, t, H, ring_cap, window, full_chunk, pe_wrap, pos_rows;
    long long ring_slot_stride;   // elements between sessions of one layer: 2*H*ring_cap*64
};
// qkv (n*t, 3*D) activation type; ring = this layer's (slot, 2, H, ring_cap, 64); ptab (pos_rows, D);
// out (n*t, D).  Appends the new K/V rows to the ring.
template <typename TA>
int attention_stream(const AttnStream& a, const TA* qkv, TA* ring, const TA* ptab, const float* pos_u,
                     const float* pos_v, TA* out, cudaStream_t st);
// offline: qkv (B*T, 3*D); valid lengths ilens (B); window from (chunk, left); positions 0..T-1.
template <typename TA>
int attention_offline(const TA* qkv, int B, int T, int H, const int32_t* ilens, int chunk, int left,
                      const TA* ptab, const float* pos_u, const float* pos_v, TA* out, cudaStream_t st);
// end of a streaming step: n_frames += t, pe_index = pe_index % wrap + chunk_size (attention.py:107,120),
// and flip the live half of the double-buffered adapter cache.  Either group may be null.
int advance_sessions(const int32_t* ids, int n, int t, int chunk_size, int pe_wrap, int32_t* n_frames,
                     int32_t* pe_index, int32_t* adapter_valid, cudaStream_t st);

// ---- conversions -----------------------------------------------------------------------------------
int f32_to_bf16(const float* src, bf16* dst, long long n, cudaStream_t st);
int bf16_to_f32(const bf16* src, float* dst, long long n, cudaStream_t st);
// dst[n][perm(k)] = src[n][k] style weight repacks (done once at finalize)
template <typename TW>
int repack_conv2(const float* w, int C, TW* out, cudaStream_t st);          // (co,ci,3,3) -> [co][(kh*3+kw)*C+ci]
template <typename TW>
int repack_sublinear(const float* w, int C, int F2, TW* out, cudaStream_t st); // [n][c*F2+f] -> [n][f*C+c]
template <typename TW>
int repack_adapter_conv(const float* w, int C2, int C, int k, TW* out, cudaStream_t st); // (co,ci,k) -> [co][tau*C+ci]
template <typename TW>
int convert_weight(const float* w, long long n, TW* out, cudaStream_t st);
// ring <-> reference (H, n, 64) layout
template <typename TA>
int ring_export(const TA* ring_kv, int H, int ring_cap, long long first_frame, int n, float* out, cudaStream_t st);
template <typename TA>
int ring_import(TA* ring_kv, int H, int ring_cap, long long first_frame, int n, const float* in, cudaStream_t st);

}  // namespace fo
