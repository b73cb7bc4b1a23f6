// FFMA GEMM: C[M,N] = A[M,K] * W[N,K]^T with the shared epilogue.  This is the fp32 parity path
// (north_star: 1e-4 max-abs in fp32 needs true fp32 products, not TF32) and the fallback for bf16
// shapes the tcgen05 kernel does not take.  64x64x16 tiles, 256 threads, 4x4 outputs per thread,
// register-prefetched global loads.  The A operand goes through AGather / RowMap (fo_common.cuh) so the
// subsampling conv2 and the adapter conv run as implicit GEMMs without materialising im2col.
#include "fo_common.cuh"

namespace fo {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct Vec4<__half> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<__half2*>(&t.x));
        const float2 b = __half22float2(*reinterpret_cast<__half2*>(&t.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
};

template <typename TIn, typename TWt, typename TAct>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TIn* __restrict__ A, AGather ga, const TWt* __restrict__ W, int M, int N, int K,
                 Epilogue ep, RowMap rmap) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Ws[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // loader mapping: one 4-wide k slice of one row per thread
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int arow = m0 + lrow, wrow = n0 + lrow;
    const bool a_ok = arow < M, w_ok = wrow < N;
    const TWt* wp = W + (long long)(w_ok ? wrow : 0) * K + lk;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[4], rw[4];
    auto fetch = [&](int k0) {
        const int seg = k0 / ga.seg_len, kin = k0 - seg * ga.seg_len;
        const long long r = (long long)arow + ga.rowoff[seg];
        if (a_ok && r < ga.rows) {
            const TIn* ap = A + ((long long)ga.plane[seg] * ga.rows + r) * ga.seg_len + kin + lk;
            Vec4<TIn>::load(ap, ra);
        } else {
            ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
        }
        if (w_ok) Vec4<TWt>::load(wp + k0, rw);
        else rw[0] = rw[1] = rw[2] = rw[3] = 0.f;
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[lk + j][lrow] = ra[j];
            Ws[lk + j][lrow] = rw[j];
        }
        __syncthreads();
        if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int nbase = n0 + tx * 4;
    if (nbase >= N) return;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = ep.bias[nbase + j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        long long drow;
        if (m >= M || !row_map(rmap, m, drow)) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = (acc[i][j] + bv[j]) * ep.scale;
            if (ep.relu) v[j] = fmaxf(v[j], 0.f);
        }
        long long o = drow * ep.ldc + nbase;
        if (ep.residual) {
            float4 r = *reinterpret_cast<const float4*>(ep.residual + o);
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        const bool wf = ep.c_f32 && (!ep.split_col || nbase < ep.split_col);
        const bool wa = ep.c_act && (!ep.split_col || nbase >= ep.split_col);
        if (wf) *reinterpret_cast<float4*>(ep.c_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
        if (wa) {
            TAct* c = reinterpret_cast<TAct*>(ep.c_act) + o;
            if (sizeof(TAct) == 2) count_sat4(ep.sat, v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = from_f<TAct>(v[j]);
        }
    }
}

}  // namespace

template <typename TIn, typename TWt>
int gemm_simt(const TIn* A, const AGather& ga, const TWt* W, int M, int N, int K, const Epilogue& ep,
              const RowMap& rmap, cudaStream_t st) {
    if (M <= 0) return 0;
    FO_CHECK(K % BK == 0 && ga.seg_len % BK == 0, "gemm_simt: K (%d) and segment (%d) must be multiples of %d", K,
             ga.seg_len, BK);
    FO_CHECK(N % 4 == 0 && ep.ldc % 4 == 0, "gemm_simt: N and ldc must be multiples of 4");
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    gemm_simt_kernel<TIn, TWt, TIn><<<grid, 256, 0, st>>>(A, ga, W, M, N, K, ep, rmap);
    FO_LAUNCHED();
    FO_CUDA(cudaGetLastError());
    return 0;
}

template int gemm_simt<float, float>(const float*, const AGather&, const float*, int, int, int, const Epilogue&, const RowMap&, cudaStream_t);
template int gemm_simt<bf16, bf16>(const bf16*, const AGather&, const bf16*, int, int, int, const Epilogue&, const RowMap&, cudaStream_t);
template int gemm_simt<__half, __half>(const __half*, const AGather&, const __half*, int, int, int, const Epilogue&, const RowMap&, cudaStream_t);

}  // namespace fo
