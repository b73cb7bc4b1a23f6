// Persistent kernel for the transformer stack of one streaming step (fp16 activation context, plain feed-forward):
// all 24 layers of  norm1 -> QKV -> chunk attention over the KV ring -> out-proj (+residual) -> norm2 -> FFN1 -> FFN2
// (+residual), then after_norm  (reference: encoder/transformer.py layer loop, encoder/attention.py:407-459,
// encoder/encoder_layer... see DESIGN.md 4c) run inside ONE cooperative launch of one CTA per SM.
//
// Why: at 64 sessions every kernel of the per-layer chain is 5-12 us of fixed cost (launch gap, barrier/TMEM set-up,
// pipeline fill from cold weights, epilogue drain) around ~1 us of HBM traffic.  Here the phases of a layer are
// separated by a grid barrier (one atomic + one polled flag, ~0.5 us) instead of a kernel boundary, and the WEIGHT
// operand never stops streaming: its TMA producer warp walks the static (layer, GEMM, unit) schedule of its CTA on its
// own, across the barriers, as far ahead as the shared-memory ring allows.
//
// CTA = 31 warps:
//   warp 0       weight producer: cp.async.bulk.tensor of 128 x 64 weight boxes (tensor maps in global memory)
//   warp 1       activation producer: NT x 64 boxes of h / att / ffh, gated by the phase barriers; its lane 0 is
//                also the CTA's delegate in the grid barrier
//   warp 2       TMEM allocation + tcgen05.mma issue (weights on the 128-row UMMA-M side, a chunk of NT <= 64 tokens
//                on the UMMA-N side, four 64-column accumulators in flight)
//   warps 3..30  28 worker warps: LayerNorm rows (one row per warp), GEMM epilogues (workers 0..7: TMEM -> registers ->
//                global, each thread owns one output column, so every store of a warp is one contiguous line), and the
//                attention items (7 groups of 4 warps, one (session, head) per group at a time, same mma.sync body as
//                attention_stream_mma_kernel)
// GEMM unit = (128 output columns, token chunk, k range).  QKV and FFN1 run full K per unit and write bias(+ReLU)'d
// fp32/fp16 outputs; out-proj and FFN2 are split along K, write raw fp32 partials [split][token][D], and the LayerNorm
// phase that follows folds  x += bias + sum_s partial_s  (fixed order: deterministic) into its row pass.
// The attention groups reuse the GEMM ring's shared memory; the weight producer therefore stops in front of the
// out-proj weights until the CTA's groups have finished the layer (mbarrier), after asking L2 for those weights.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fo_common.cuh"

namespace fo {

int tc_make_map(CUtensorMap* map, const void* base, int seg_len, long long rows, int planes, int box_rows);

namespace {

constexpr int DK = 64;
constexpr int ST_GROUPS = 7;                     // attention groups of 4 warps
constexpr int ST_WORKERS = ST_GROUPS * 4;
constexpr int ST_THREADS = (3 + ST_WORKERS) * 32;   // 992
constexpr int ST_PART = ST_THREADS - 32;         // threads that meet at the phase barriers (all but the weight producer)
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int ST_STAGES = 8;
constexpr int ST_NT = 64;                        // token chunk = UMMA N
constexpr int ST_NACC = 4;                       // TMEM accumulators (64 columns each)
constexpr int ST_EPI = 8;                        // epilogue warps (workers 0..7)
constexpr uint32_t W_BYTES = BM * BK * 2;        // 16 KB
constexpr uint32_t A_BYTES = ST_NT * BK * 2;     // 8 KB slot (NT may be smaller)
constexpr uint32_t STAGE_BYTES = W_BYTES + A_BYTES;
constexpr uint32_t RING_BYTES = ST_STAGES * STAGE_BYTES;     // 192 KB
constexpr uint32_t SMEM_MAX = 226 * 1024;

struct StackLayer {
    const float *ln1g, *ln1b, *ln2g, *ln2b, *bqkv, *bo, *b1, *b2, *pos_u, *pos_v;
    const __half* ptab_h;
    __half* ring;                 // this layer's (slot, 2, H, ring_cap, 64)
};

struct StackParams {
    const StackLayer* layers;
    const CUtensorMap* wmaps;     // [L][4]: wqkv, wo, w1, w2 (box 128 x 64)
    int L, D, FF, H, M, NT, chunks;
    int split_o, split_f2;
    float* x;                     // residual stream (M, D) fp32
    __half* h;                    // LayerNorm output (M, D)
    __half* qkv;                  // (M, 3D): K | V columns used
    float* q32;                   // (M, 3D): Q columns used
    __half* att;                  // (M, D)
    __half* ffh;                  // (M, FF)
    float* partial;               // [split][M][D]
    const float *fin_g, *fin_b;
    float* enc_out;               // after_norm output (M, D) fp32
    AttnStream a;
    unsigned int* gbar;           // grid barrier counter (zeroed before the launch)
    int att_groups, att_group_bytes;
    float eps;
    int dbg;                      // development (FO_STACK_DBG): 1 skip LayerNorm, 2 skip GEMMs, 4 skip attention (timing only)
    unsigned long long* trace;    // development (FO_STACK_TRACE=1): %globaltimer of CTA 0 after every grid barrier
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// Every wait of this kernel is bounded: a protocol bug traps (the launch fails) instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try(bar, parity)) return;
        if ((spin & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
constexpr uint64_t UMMA_DESC_HI = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) { return UMMA_DESC_HI | (uint64_t)((saddr & 0x3FFFFu) >> 4); }

__device__ __forceinline__ void part_bar() { asm volatile("bar.sync 1, %0;" ::"n"(ST_PART) : "memory"); }
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }

// ---- grid barrier -----------------------------------------------------------------------------------
// Monotonic counter: generation g is complete when the counter reaches g * gridDim.x.  Called by the ST_PART threads.
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int& gen, unsigned long long* trace = nullptr) {
    part_bar();
    ++gen;
    if (threadIdx.x == 32) {
        const unsigned int target = gen * gridDim.x;
        // release: the CTA's stores (ordered before this thread by the barrier above) are visible before the count moves
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        long long t0 = 0;
        for (uint32_t spin = 0;; ++spin) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            if ((spin & 255u) == 255u) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ll) __trap();
            }
        }
        asm volatile("fence.proxy.async;" ::: "memory");      // other CTAs' generic stores before this thread's TMA loads
        if (trace && blockIdx.x == 0) trace[gen] = gtime();
    }
    part_bar();
}

// ---- GEMM unit schedule -------------------------------------------------------------------------------
struct GemmDesc {
    int tiles, kblocks, split, kbs, units;
};
__device__ __forceinline__ GemmDesc gemm_desc(const StackParams& p, int g) {
    GemmDesc d;
    d.tiles = (g == 0 ? 3 * p.D : g == 2 ? p.FF : p.D) / BM;
    d.kblocks = (g == 3 ? p.FF : p.D) / BK;
    d.split = g == 1 ? p.split_o : g == 3 ? p.split_f2 : 1;
    d.kbs = d.kblocks / d.split;
    d.units = d.tiles * p.chunks * d.split;
    return d;
}
// unit -> (tile, chunk, split): tiles fastest, so neighbouring CTAs share the activation chunk
__device__ __forceinline__ void unit_decode(const GemmDesc& d, int chunks, int u, int& tile, int& chunk, int& sp) {
    tile = u % d.tiles;
    const int r = u / d.tiles;
    chunk = r % chunks;
    sp = r / chunks;
}

struct RingPos {          // position in the shared-memory stage ring, kept in lock-step by both producers and the MMA warp
    int s;
    uint32_t ph;
    __device__ __forceinline__ void advance() { if (++s == ST_STAGES) { s = 0; ph ^= 1u; } }
};

// LayerNorm phase: one row per 4-warp group (lane = two float4 of the row, so that x, the bias and every split-K partial
// of a row are in flight together: with one warp per row the 64-register budget serialises the loads).
// x += bias + sum_s partial_s (when nsplit > 0, fixed order), then the row norm.
__device__ __noinline__ void ln_phase(const StackParams& p, int gi, int gtid, float* red, int nsplit,
                                      const float* __restrict__ bias, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, __half* __restrict__ out16, float* __restrict__ out32) {
    const int D = p.D, nv = D >> 9;                    // float4 per thread (D = 512 or 1024)
    const long long MD4 = ((long long)p.M * D) >> 2;
    const int warp = gtid >> 5, lane = gtid & 31;
    for (int r = blockIdx.x + gridDim.x * gi; r < p.M; r += gridDim.x * ST_GROUPS) {
        float4* xr = reinterpret_cast<float4*>(p.x + (long long)r * D) + gtid;
        float4 v[2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < nv) v[i] = __ldcg(xr + i * 128);
        if (nsplit > 0) {
            const float4* pr = reinterpret_cast<const float4*>(p.partial) + (long long)r * (D >> 2) + gtid;
            float4 b[2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < nv) b[i] = *(reinterpret_cast<const float4*>(bias) + gtid + i * 128);
            for (int s = 0; s < nsplit; ++s) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (i < nv) {
                        const float4 q = __ldcg(pr + s * MD4 + i * 128);
                        b[i].x += q.x; b[i].y += q.y; b[i].z += q.z; b[i].w += q.w;
                    }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < nv) {
                    v[i].x += b[i].x; v[i].y += b[i].y; v[i].z += b[i].z; v[i].w += b[i].w;
                    xr[i * 128] = v[i];
                }
        }
        float4 g[2], bt[2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < nv) {
                g[i] = *(reinterpret_cast<const float4*>(gamma) + gtid + i * 128);
                bt[i] = *(reinterpret_cast<const float4*>(beta) + gtid + i * 128);
            }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < nv) sum += v[i].x + v[i].y + v[i].z + v[i].w;
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;
        group_bar(gi);
        const float mu = (red[0] + red[1] + red[2] + red[3]) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < nv) {
                const float a0 = v[i].x - mu, a1 = v[i].y - mu, a2 = v[i].z - mu, a3 = v[i].w - mu;
                q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            }
        q = warp_sum(q);
        if (lane == 0) red[4 + warp] = q;
        group_bar(gi);
        const float rstd = rsqrtf((red[4] + red[5] + red[6] + red[7]) / D + p.eps);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i < nv) {
                const float o0 = (v[i].x - mu) * rstd * g[i].x + bt[i].x, o1 = (v[i].y - mu) * rstd * g[i].y + bt[i].y;
                const float o2 = (v[i].z - mu) * rstd * g[i].z + bt[i].z, o3 = (v[i].w - mu) * rstd * g[i].w + bt[i].w;
                const long long off = (long long)r * D + (gtid + i * 128) * 4;
                if (out32) *reinterpret_cast<float4*>(out32 + off) = make_float4(o0, o1, o2, o3);
                if (out16) {
                    uint2 hh;
                    hh.x = pack2<__half>(o0, o1);
                    hh.y = pack2<__half>(o2, o3);
                    *reinterpret_cast<uint2*>(out16 + off) = hh;
                }
            }
        group_bar(gi);                                 // red[] is reused by the next row
    }
}

// One (session, head) of the streaming attention by a group of 4 warps; the body of attention_stream_mma_kernel with the
// CTA barrier replaced by the group's named barrier and cross-SM inputs read through L2 (ld.global.cg).
__device__ __noinline__ void attn_item(const StackParams& p, const StackLayer& ly, int b, int h, unsigned char* smem,
                                          uint32_t bar, uint32_t bar_parity, int gi, int tid) {
    typedef __half TA;
    constexpr int EPC = 8, NCH = 8, GT = 128;
    const AttnStream& a = p.a;
    const int cap = a.ring_cap;
    const int t = a.t, D = a.H * DK;
    const int rows = a.window + t;
    const int vrows = (rows + 15) & ~15;
    TA* Ks = reinterpret_cast<TA*>(smem);
    TA* Ps = Ks + rows * DK;
    TA* Vs = Ps + rows * DK;
    TA* quh = Vs + vrows * DK;
    TA* qvh = quh + t * DK;
    TA* ph = qvh + t * DK;
    float* sc = reinterpret_cast<float*>(ph + t * vrows);

    const int slot = a.ids[b];
    const int nf = a.n_frames[slot];
    const int cl = min(nf, a.window);
    const int first = nf - cl;
    const int nk = cl + t;
    const int pe = a.pe_index[slot] % a.pe_wrap;
    const int start = max(0, pe - a.full_chunk);
    TA* ringK = ly.ring + (long long)slot * a.ring_slot_stride + (long long)h * cap * DK;
    TA* ringV = ringK + (long long)a.H * cap * DK;
    const TA* ptab_h = ly.ptab_h;
    const int warp = tid >> 5, lane = tid & 31;
    const int np = min(nk, a.pos_rows - start);

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses of the last item before the bulk copies
    group_bar(gi);                                   // the previous item of this group is fully done with the buffers
    if (tid == 0) {
        const int p0 = first % cap;
        const int len1 = min(cl, cap - p0), len2 = cl - len1;
        const uint32_t rowb = DK * sizeof(TA);
        mbar_expect_tx(bar, (2u * cl + np) * rowb);
        bulk_g2s(smem_u32(Ps), ptab_h + ((long long)h * a.pos_rows + start) * DK, np * rowb, bar);
        if (len1 > 0) {
            bulk_g2s(smem_u32(Ks), ringK + (long long)p0 * DK, len1 * rowb, bar);
            bulk_g2s(smem_u32(Vs), ringV + (long long)p0 * DK, len1 * rowb, bar);
        }
        if (len2 > 0) {
            bulk_g2s(smem_u32(Ks + len1 * DK), ringK, len2 * rowb, bar);
            bulk_g2s(smem_u32(Vs + len1 * DK), ringV, len2 * rowb, bar);
        }
    }
    const float* q32 = p.q32;
    const TA* qkv = p.qkv;
    for (int i = tid; i < t * NCH * 2; i += GT) {
        const int which = i / (t * NCH), r = (i / NCH) % t, c = i % NCH;
        const uint4 val = __ldcg(reinterpret_cast<const uint4*>(qkv + (long long)(b * t + r) * 3 * D + (which + 1) * D + h * DK + c * EPC));
        const int pc = c ^ ((nf + r) & 7);
        *reinterpret_cast<uint4*>((which ? Vs : Ks) + (cl + r) * DK + pc * EPC) = val;
        *reinterpret_cast<uint4*>((which ? ringV : ringK) + (long long)((nf + r) % cap) * DK + pc * EPC) = val;
    }
    for (int i = tid; i < (vrows - nk) * NCH; i += GT)
        *reinterpret_cast<uint4*>(Vs + (nk + i / NCH) * DK + (i % NCH) * EPC) = make_uint4(0, 0, 0, 0);
    for (int i = tid + np * NCH; i < nk * NCH; i += GT) {
        const int j = i / NCH, c = i % NCH;
        *reinterpret_cast<uint4*>(Ps + j * DK + (c ^ ((start + j) & 7)) * EPC) =
            *reinterpret_cast<const uint4*>(ptab_h + ((long long)h * a.pos_rows + a.pos_rows - 1) * DK + (c ^ ((a.pos_rows - 1) & 7)) * EPC);
    }
    for (int i = tid; i < t * (DK / 4); i += GT) {
        const int r = i / (DK / 4), d = (i % (DK / 4)) * 4;
        const float4 q = __ldcg(reinterpret_cast<const float4*>(q32 + (long long)(b * t + r) * 3 * D + h * DK + d));
        const float4 u = *reinterpret_cast<const float4*>(ly.pos_u + h * DK + d);
        const float4 v = *reinterpret_cast<const float4*>(ly.pos_v + h * DK + d);
        const int o = r * DK + (((d >> 3) ^ (r & 7)) << 3) + (d & 7);
        uint2 hu, hv;
        hu.x = pack2<TA>(q.x + u.x, q.y + u.y); hu.y = pack2<TA>(q.z + u.z, q.w + u.w);
        hv.x = pack2<TA>(q.x + v.x, q.y + v.y); hv.y = pack2<TA>(q.z + v.z, q.w + v.w);
        *reinterpret_cast<uint2*>(quh + o) = hu;
        *reinterpret_cast<uint2*>(qvh + o) = hv;
    }
    group_bar(gi);
    mbar_wait(bar, bar_parity);

    const int g = lane >> 2, c = lane & 3;
    for (int k0 = warp * 16; k0 < nk; k0 += 64) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        const int kk0 = (first + k0 + g) & 7, kk1 = (first + k0 + g + 8) & 7;
        const int kp0 = (start + k0 + g) & 7, kp1 = (start + k0 + g + 8) & 7;
        const TA* kr0 = Ks + (k0 + g) * DK + c * 2;
        const TA* kr1 = Ks + (k0 + g + 8) * DK + c * 2;
        const TA* pr0 = Ps + (k0 + g) * DK + c * 2;
        const TA* pr1 = Ps + (k0 + g + 8) * DK + c * 2;
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks) {
            const int c0 = ks * 2, c1 = ks * 2 + 1;
            uint32_t af[4], bf[2];
            af[0] = lds32(kr0 + ((c0 ^ kk0) << 3));
            af[1] = lds32(kr1 + ((c0 ^ kk1) << 3));
            af[2] = lds32(kr0 + ((c1 ^ kk0) << 3));
            af[3] = lds32(kr1 + ((c1 ^ kk1) << 3));
            bf[0] = g < t ? lds32(quh + g * DK + ((c0 ^ g) << 3) + c * 2) : 0u;
            bf[1] = g < t ? lds32(quh + g * DK + ((c1 ^ g) << 3) + c * 2) : 0u;
            mma16816(d, af, bf);
            af[0] = lds32(pr0 + ((c0 ^ kp0) << 3));
            af[1] = lds32(pr1 + ((c0 ^ kp1) << 3));
            af[2] = lds32(pr0 + ((c1 ^ kp0) << 3));
            af[3] = lds32(pr1 + ((c1 ^ kp1) << 3));
            bf[0] = g < t ? lds32(qvh + g * DK + ((c0 ^ g) << 3) + c * 2) : 0u;
            bf[1] = g < t ? lds32(qvh + g * DK + ((c1 ^ g) << 3) + c * 2) : 0u;
            mma16816(d, af, bf);
        }
        const int q0 = c * 2;
        if (q0 < t) { sc[q0 * vrows + k0 + g] = d[0] * 0.125f; sc[q0 * vrows + k0 + g + 8] = d[2] * 0.125f; }
        if (q0 + 1 < t) { sc[(q0 + 1) * vrows + k0 + g] = d[1] * 0.125f; sc[(q0 + 1) * vrows + k0 + g + 8] = d[3] * 0.125f; }
    }
    group_bar(gi);
    for (int i = warp; i < t; i += 4) {
        float sv[4];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = lane + 32 * k;
            sv[k] = j < nk ? sc[i * vrows + j] : -INFINITY;
            m = fmaxf(m, sv[k]);
        }
        m = warp_max(m);
        float ssum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sv[k] = (lane + 32 * k < nk) ? __expf(sv[k] - m) : 0.f;
            ssum += sv[k];
        }
        ssum = warp_sum(ssum);
        const float inv = 1.f / ssum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = lane + 32 * k;
            if (j < vrows) ph[i * vrows + j] = __float2half_rn(sv[k] * inv);
        }
    }
    group_bar(gi);
    {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        const int dim0 = warp * 16;
        const int mi = lane >> 3, r = lane & 7;
        for (int key0 = 0; key0 < nk; key0 += 16) {
            uint32_t af[4], bf[2];
            const int vrow = key0 + r + ((mi & 2) ? 8 : 0);
            const TA* ap = Vs + vrow * DK + ((((dim0 >> 3) + (mi & 1)) ^ ((first + vrow) & 7)) << 3);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3])
                         : "r"(smem_u32(ap)));
            bf[0] = g < t ? lds32(ph + g * vrows + key0 + c * 2) : 0u;
            bf[1] = g < t ? lds32(ph + g * vrows + key0 + 8 + c * 2) : 0u;
            mma16816(d, af, bf);
        }
        const int q0 = c * 2;
        if (q0 < t) { sc[q0 * DK + dim0 + g] = d[0]; sc[q0 * DK + dim0 + g + 8] = d[2]; }
        if (q0 + 1 < t) { sc[(q0 + 1) * DK + dim0 + g] = d[1]; sc[(q0 + 1) * DK + dim0 + g + 8] = d[3]; }
    }
    group_bar(gi);
    for (int i = tid; i < t * (DK / 2); i += GT) {
        const int q = i / (DK / 2), pr = i % (DK / 2);
        *reinterpret_cast<uint32_t*>(p.att + (long long)(b * t + q) * D + h * DK + 2 * pr) =
            pack2<__half>(sc[q * DK + 2 * pr], sc[q * DK + 2 * pr + 1]);
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS, 1)
stack_stream_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_att,
                    const __grid_constant__ CUtensorMap map_ffh, const StackParams p) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_w[ST_STAGES];
    __shared__ __align__(8) uint64_t full_a[ST_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[ST_STAGES];
    __shared__ __align__(8) uint64_t acc_full[ST_NACC];
    __shared__ __align__(8) uint64_t acc_empty[ST_NACC];
    __shared__ __align__(8) uint64_t attn_gate;
    __shared__ __align__(8) uint64_t attn_load[ST_GROUPS];
    __shared__ uint32_t tmem_base_s;
    __shared__ float ln_red[ST_GROUPS][8];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* smem_al = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t fw0 = smem_u32(&full_w[0]), fa0 = smem_u32(&full_a[0]), em0 = smem_u32(&empty_bar[0]);
    const uint32_t af0 = smem_u32(&acc_full[0]), ae0 = smem_u32(&acc_empty[0]), gate = smem_u32(&attn_gate);

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST_STAGES; ++s) { mbar_init(fw0 + 8 * s, 1); mbar_init(fa0 + 8 * s, 1); mbar_init(em0 + 8 * s, 1); }
        for (int i = 0; i < ST_NACC; ++i) { mbar_init(af0 + 8 * i, 1); mbar_init(ae0 + 8 * i, ST_EPI); }
        mbar_init(gate, p.att_groups);
        for (int i = 0; i < ST_GROUPS; ++i) mbar_init(smem_u32(&attn_load[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(ST_NACC * ST_NT))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_bytes = (uint32_t)p.NT * BK * 2;

    if (warp == 0) {
        // ---- weight producer: the whole step's schedule of this CTA, ahead of everything else ----
        if (lane == 0 && !(p.dbg & 2)) {
            RingPos rp{0, 1u};
            for (int l = 0; l < p.L; ++l) {
                for (int g = 0; g < 4; ++g) {
                    const GemmDesc d = gemm_desc(p, g);
                    const CUtensorMap* map = p.wmaps + l * 4 + g;
                    if (g == 1) {
                        // the attention groups own the ring's memory: ask L2 for this CTA's out-proj boxes, then wait for them
                        for (int u = cta; u < d.units; u += G) {
                            int tile, chunk, sp;
                            unit_decode(d, p.chunks, u, tile, chunk, sp);
                            for (int i = 0; i < d.kbs; ++i) tma_prefetch_3d(map, (sp * d.kbs + i) * BK, tile * BM, 0);
                        }
                        mbar_wait(gate, (uint32_t)(l & 1));
                    }
                    for (int u = cta; u < d.units; u += G) {
                        int tile, chunk, sp;
                        unit_decode(d, p.chunks, u, tile, chunk, sp);
                        const int kb0 = sp * d.kbs;
                        for (int i = 0; i < d.kbs; ++i) {
                            mbar_wait(em0 + 8 * rp.s, rp.ph);
                            mbar_expect_tx(fw0 + 8 * rp.s, W_BYTES);
                            tma_load_3d(base + rp.s * STAGE_BYTES, map, fw0 + 8 * rp.s, (kb0 + i) * BK, tile * BM, 0);
                            rp.advance();
                        }
                    }
                }
            }
        }
    } else {
        const int wi = warp - 3;                         // worker index (negative for the two role warps)
        unsigned int gen = 0;
        if (p.trace && threadIdx.x == 32 && blockIdx.x == 0) p.trace[0] = gtime();
        RingPos rp{0, warp == 1 ? 1u : 0u};
        int acc_cnt = 0;                                 // accumulators handed over so far (MMA warp and epilogue warps)
        int items_done = 0;                              // attention items of this group so far (bulk-load barrier parity)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // fp16 x fp16 -> fp32

        auto gemm_phase = [&](int l, int g) {
            if (p.dbg & 2) return;
            const GemmDesc d = gemm_desc(p, g);
            if (warp == 1) {
                if (lane == 0) {
                    const CUtensorMap* map = g <= 0 ? &map_h : g == 1 ? &map_att : g == 2 ? &map_h : &map_ffh;
                    for (int u = cta; u < d.units; u += G) {
                        int tile, chunk, sp;
                        unit_decode(d, p.chunks, u, tile, chunk, sp);
                        const int kb0 = sp * d.kbs;
                        for (int i = 0; i < d.kbs; ++i) {
                            mbar_wait(em0 + 8 * rp.s, rp.ph);
                            mbar_expect_tx(fa0 + 8 * rp.s, a_bytes);
                            tma_load_3d(base + rp.s * STAGE_BYTES + W_BYTES, map, fa0 + 8 * rp.s, (kb0 + i) * BK, chunk * p.NT, 0);
                            rp.advance();
                        }
                    }
                }
                __syncwarp();
            } else if (warp == 2) {
                if (lane == 0) {
                    for (int u = cta; u < d.units; u += G) {
                        const int ab = acc_cnt % ST_NACC;
                        mbar_wait(ae0 + 8 * ab, (((uint32_t)(acc_cnt / ST_NACC)) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t tacc = tmem_base + (uint32_t)(ab * ST_NT);
                        for (int i = 0; i < d.kbs; ++i) {
                            mbar_wait(fw0 + 8 * rp.s, rp.ph);
                            mbar_wait(fa0 + 8 * rp.s, rp.ph);
                            tc_fence_after();
                            const uint32_t sa = base + rp.s * STAGE_BYTES;
                            const uint64_t da = umma_desc(sa), db = umma_desc(sa + W_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) tc_mma(tacc, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                            tc_commit(em0 + 8 * rp.s);
                            rp.advance();
                        }
                        tc_commit(af0 + 8 * ab);
                        ++acc_cnt;
                    }
                }
                __syncwarp();
            } else if (wi == ST_EPI) {
                // an idle worker asks L2 for the weight boxes of this CTA's units of the NEXT GEMM (the producer's own
                // run-ahead is 8 stages; HBM latency is what the first k-blocks of a phase would otherwise wait for)
                const int g2 = (g + 1) & 3, l2 = g == 3 ? l + 1 : l;
                if (l2 < p.L) {
                    const GemmDesc d2 = gemm_desc(p, g2);
                    const CUtensorMap* map2 = p.wmaps + l2 * 4 + g2;
                    for (int u = cta; u < d2.units; u += G) {
                        int tile, chunk, sp;
                        unit_decode(d2, p.chunks, u, tile, chunk, sp);
                        for (int i = lane; i < d2.kbs; i += 32) tma_prefetch_3d(map2, (sp * d2.kbs + i) * BK, tile * BM, 0);
                    }
                }
            } else if (wi < ST_EPI) {
                const StackLayer& ly = p.layers[l];
                const int q = warp & 3;                  // TMEM lane quarter this warp may read
                const int half = wi >> 2;
                const float* bias = g == 0 ? ly.bqkv : g == 2 ? ly.b1 : nullptr;
                for (int u = cta; u < d.units; u += G) {
                    int tile, chunk, sp;
                    unit_decode(d, p.chunks, u, tile, chunk, sp);
                    const int ab = acc_cnt % ST_NACC;
                    mbar_wait(af0 + 8 * ab, ((uint32_t)(acc_cnt / ST_NACC)) & 1u);
                    tc_fence_after();
                    const int n = tile * BM + q * 32 + lane;             // output column of this thread
                    const float bv = bias ? bias[n] : 0.f;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * ST_NT);
                    for (int c16 = half; c16 * 16 < p.NT; c16 += 2) {
                        float v[16];
                        tc_ld16(taddr + c16 * 16, v);
                        const int m0 = chunk * p.NT + c16 * 16;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int m = m0 + j;
                            if (m < p.M) {
                                if (g == 0) {
                                    const float o = v[j] + bv;
                                    if (n < p.D) p.q32[(long long)m * 3 * p.D + n] = o;
                                    else p.qkv[(long long)m * 3 * p.D + n] = from_f<__half>(o);
                                } else if (g == 2) {
                                    p.ffh[(long long)m * p.FF + n] = from_f<__half>(fmaxf(v[j] + bv, 0.f));
                                } else {
                                    __stcg(p.partial + ((long long)sp * p.M + m) * p.D + n, v[j]);
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ae0 + 8 * ab);
                    ++acc_cnt;
                }
            }
        };

        for (int l = 0; l < p.L; ++l) {
            const StackLayer& ly = p.layers[l];
            // norm1 (folds the previous layer's FFN2 partials into the residual stream)
            if (wi >= 0 && !(p.dbg & 1)) ln_phase(p, wi >> 2, (wi & 3) * 32 + lane, ln_red[wi >> 2], l == 0 ? 0 : p.split_f2, l == 0 ? nullptr : p.layers[l - 1].b2, ly.ln1g, ly.ln1b, p.h, nullptr);
            grid_barrier(p.gbar, gen, p.trace);
            gemm_phase(l, 0);
            grid_barrier(p.gbar, gen, p.trace);
            // attention: items (session, head) of this CTA over its groups
            if (wi >= 0) {
                const int gi = wi >> 2;
                if (gi < p.att_groups) {
                    const int gtid = (wi & 3) * 32 + lane;
                    const int n_items = p.a.n * p.H;
                    for (int it = cta + gi * G; it < n_items && !(p.dbg & 4); it += p.att_groups * G) {
                        attn_item(p, ly, it / p.H, it % p.H, smem_al + (size_t)gi * p.att_group_bytes, smem_u32(&attn_load[gi]),
                                  (uint32_t)(items_done & 1), gi, gtid);
                        ++items_done;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring's next writer is TMA
                    group_bar(gi);
                    if (gtid == 0) mbar_arrive(gate);
                }
            }
            grid_barrier(p.gbar, gen, p.trace);
            gemm_phase(l, 1);
            grid_barrier(p.gbar, gen, p.trace);
            if (wi >= 0 && !(p.dbg & 1)) ln_phase(p, wi >> 2, (wi & 3) * 32 + lane, ln_red[wi >> 2], p.split_o, ly.bo, ly.ln2g, ly.ln2b, p.h, nullptr);
            grid_barrier(p.gbar, gen, p.trace);
            gemm_phase(l, 2);
            grid_barrier(p.gbar, gen, p.trace);
            gemm_phase(l, 3);
            grid_barrier(p.gbar, gen, p.trace);
        }
        if (wi >= 0) ln_phase(p, wi >> 2, (wi & 3) * 32 + lane, ln_red[wi >> 2], p.split_f2, p.layers[p.L - 1].b2, p.fin_g, p.fin_b, nullptr, p.enc_out);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(ST_NACC * ST_NT))
                     : "memory");
    }
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------
struct StackState {
    StackLayer* layers_dev = nullptr;
    CUtensorMap* wmaps_dev = nullptr;
    unsigned int* gbar = nullptr;
    int sm_count = 0;
    bool attr_set = false;
};

int stack_state_create(StackState** out) {
    *out = new StackState();
    return 0;
}
void stack_state_destroy(StackState* s) {
    if (!s) return;
    cudaFree(s->layers_dev);
    cudaFree(s->wmaps_dev);
    cudaFree(s->gbar);
    delete s;
}

int stack_stream_launch(StackState* s, const StackHostArgs& ha, cudaStream_t st) {
    const int D = ha.D, FF = ha.FF, H = ha.H, L = ha.L, M = ha.a.n * ha.a.t;
    // shapes this kernel covers; anything else runs the per-kernel chain
    if (D % 512 != 0 || D > 1024 || D != H * DK || FF % BM != 0 || (3 * D) % BM != 0 || ha.a.t > 8 || M <= 0) return 1;
    if (!s->layers_dev) {
        int dev = 0;
        FO_CUDA(cudaGetDevice(&dev));
        FO_CUDA(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev));
        std::vector<StackLayer> hl(L);
        std::vector<CUtensorMap> hm((size_t)L * 4);
        for (int l = 0; l < L; ++l) {
            const StackLayerHost& w = ha.layers[l];
            hl[l] = StackLayer{w.ln1g, w.ln1b, w.ln2g, w.ln2b, w.bqkv, w.bo, w.b1, w.b2, w.pos_u, w.pos_v,
                               reinterpret_cast<const __half*>(w.ptab_h), reinterpret_cast<__half*>(w.ring)};
            FO_TRY(tc_make_map(&hm[l * 4 + 0], w.wqkv, D, 3 * D, 1, BM));
            FO_TRY(tc_make_map(&hm[l * 4 + 1], w.wo, D, D, 1, BM));
            FO_TRY(tc_make_map(&hm[l * 4 + 2], w.w1, D, FF, 1, BM));
            FO_TRY(tc_make_map(&hm[l * 4 + 3], w.w2, FF, D, 1, BM));
        }
        FO_CUDA(cudaMalloc(&s->layers_dev, sizeof(StackLayer) * L));
        FO_CUDA(cudaMalloc(&s->wmaps_dev, sizeof(CUtensorMap) * L * 4));
        FO_CUDA(cudaMalloc(&s->gbar, 256));
        FO_CUDA(cudaMemcpy(s->layers_dev, hl.data(), sizeof(StackLayer) * L, cudaMemcpyHostToDevice));
        FO_CUDA(cudaMemcpy(s->wmaps_dev, hm.data(), sizeof(CUtensorMap) * L * 4, cudaMemcpyHostToDevice));
    }
    StackParams p;
    memset(&p, 0, sizeof(p));
    p.layers = s->layers_dev;
    p.wmaps = s->wmaps_dev;
    p.L = L; p.D = D; p.FF = FF; p.H = H; p.M = M;
    p.NT = M >= ST_NT ? ST_NT : (M + 15) / 16 * 16;
    p.chunks = (M + p.NT - 1) / p.NT;
    // K splits of the two residual GEMMs: ~4 k-blocks per unit, as many units as fit one wave
    auto pick_split = [&](int tiles, int kblocks) {
        int sp = 1;
        while (sp * 2 <= ha.max_split && kblocks % (sp * 2) == 0 && kblocks / (sp * 2) >= 4 && tiles * p.chunks * sp * 2 <= s->sm_count) sp *= 2;
        return sp;
    };
    p.split_o = pick_split(D / BM, D / BK);
    p.split_f2 = pick_split(D / BM, FF / BK);
    if (ha.force_split_o > 0) p.split_o = ha.force_split_o;
    if (ha.force_split_f2 > 0) p.split_f2 = ha.force_split_f2;
    if ((D / BK) % p.split_o != 0 || (FF / BK) % p.split_f2 != 0 || p.split_o > ha.max_split || p.split_f2 > ha.max_split) return 1;
    p.x = ha.x; p.h = reinterpret_cast<__half*>(ha.h); p.qkv = reinterpret_cast<__half*>(ha.qkv); p.q32 = ha.q32;
    p.att = reinterpret_cast<__half*>(ha.att); p.ffh = reinterpret_cast<__half*>(ha.ffh); p.partial = ha.partial;
    p.fin_g = ha.fin_g; p.fin_b = ha.fin_b; p.enc_out = ha.enc_out;
    p.a = ha.a;
    p.gbar = s->gbar;
    p.eps = 1e-5f;
    const int rows = ha.a.window + ha.a.t, vrows = (rows + 15) & ~15;
    const int gbytes = (2 * rows + vrows) * DK * 2 + 2 * ha.a.t * DK * 2 + ha.a.t * vrows * 2 + ha.a.t * std::max(vrows, DK) * 4;
    p.att_group_bytes = (gbytes + 127) / 128 * 128;
    if (rows > 128) return 1;                                   // softmax strip covers 128 keys
    p.att_groups = std::min<int>(ST_GROUPS, (int)((SMEM_MAX - 1024) / p.att_group_bytes));
    if (p.att_groups < 1) return 1;
    const size_t smem = std::max<size_t>(RING_BYTES, (size_t)p.att_groups * p.att_group_bytes) + 1024;
    if (smem > SMEM_MAX) return 1;
    if (!s->attr_set) {
        FO_CUDA(cudaFuncSetAttribute(stack_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        s->attr_set = true;
    }
    CUtensorMap mh, ma, mf;
    FO_TRY(tc_make_map(&mh, ha.h, D, M, 1, p.NT));
    FO_TRY(tc_make_map(&ma, ha.att, D, M, 1, p.NT));
    FO_TRY(tc_make_map(&mf, ha.ffh, FF, M, 1, p.NT));
    FO_CUDA(cudaMemsetAsync(s->gbar, 0, 4, st));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(s->sm_count);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int trace_on = -1;
    static unsigned long long* trace_buf = nullptr;
    if (trace_on < 0) {
        const char* e = getenv("FO_STACK_TRACE");
        trace_on = (e && e[0] == '1') ? 1 : 0;
        if (trace_on) cudaMalloc(&trace_buf, 1024 * sizeof(unsigned long long));
    }
    p.trace = trace_on ? trace_buf : nullptr;
    {
        const char* e = getenv("FO_STACK_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
    FO_CUDA(cudaLaunchKernelEx(&cfg, stack_stream_kernel, mh, ma, mf, p));
    FO_LAUNCHED();
    FO_CUDA(cudaGetLastError());
    if (trace_on && L * 7 + 1 < 1024) {
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> hb(L * 7 + 1);
        cudaMemcpy(hb.data(), trace_buf, hb.size() * 8, cudaMemcpyDeviceToHost);
        static const char* names[7] = {"ln1", "qkv", "attn", "out", "ln2", "ffn1", "ffn2"};
        fprintf(stderr, "stack_trace M=%d NT=%d chunks=%d split_o=%d split_f2=%d groups=%d total=%lld ns\n", M, p.NT, p.chunks, p.split_o,
                p.split_f2, p.att_groups, (long long)(hb[L * 7] - hb[0]));
        for (int l : {0, L / 2, L - 1}) {
            fprintf(stderr, "  layer %d:", l);
            for (int k = 0; k < 7; ++k) fprintf(stderr, " %s=%lld", names[k], (long long)(hb[l * 7 + k + 1] - hb[l * 7 + k]));
            fprintf(stderr, " ns\n");
        }
    }
    return 0;
}

}  // namespace fo
