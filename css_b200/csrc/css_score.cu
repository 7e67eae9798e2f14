// Stage 3: anchor / negative sampling, query x {positive prototype, negatives} scoring at temperature, cross-entropy and
// d loss / d anchor in ONE pass over the gathered rows, then the dense-gradient scatter of the backward.
// Reference: generalframeworks/loss/loss.py:124-149 (+ negative_index_sampler :410-418) and its autograd backward.
//
// Every query draws its own Nn negatives (loss.py:137-142), so this is a per-query row gather (1 KB fp32 rows of the
// pixel-major normalised copy written by css_stream_rep), not a shared-operand GEMM: 2*D flops per 4*D gathered bytes.
// It is bound by L2/HBM gather bandwidth, not by the FMA or tensor pipes.
//
// Work split: one warp per (present-class slot k, query q).  The warp is four 8-lane groups; a group owns one candidate
// row at a time (8 lanes x 8 x 128-bit loads = the 1 KB row, each 128 B line read by one 8-lane group), reduces the dot
// with 3 shuffles, and keeps an online-softmax state (m, l, sum_j e^{z_j-m} r_hat_j) in registers, so the backward never
// re-gathers.  Candidate row ids for the next 32 candidates are produced one per lane (Philox draw or fed index ->
// rotated segment -> class list lookup) and handed to the groups by shuffle.
#include "css_common.cuh"

#define SC_WARPS 4
#define SC_THREADS (SC_WARPS * 32)

struct DrawKey {
    uint2 key;       // seed
    uint32_t off_lo, off_hi;
};

__device__ __forceinline__ DrawKey make_key(uint64_t seed, uint64_t offset) {
    DrawKey k;
    k.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    k.off_lo = (uint32_t)offset;
    k.off_hi = (uint32_t)(offset >> 32);
    return k;
}

// anchor draw: uniform index into the class's hard list (loss.py:127)
__device__ __forceinline__ int draw_anchor(const DrawKey& dk, int k, int q, int n_hard) {
    const uint4 r = Philox::run(make_uint4(0xffffffffu, ((uint32_t)k << 24) | (uint32_t)q, dk.off_lo, dk.off_hi), dk.key);
    return (int)__umulhi(r.x, (uint32_t)n_hard);
}

// negative draw j in [0, Nn): class ~ Categorical(cdf row), index ~ Uniform inside that class's valid list; returned as
// an index into the rotated concatenation of the valid lists (loss.py:136-142)
__device__ __forceinline__ int draw_negative(const DrawKey& dk, int k, int q, int j, int V, const float* cdf_row,
                                             const int* rot_off) {
    const uint4 r = Philox::run(make_uint4((uint32_t)j, ((uint32_t)k << 24) | (uint32_t)q, dk.off_lo, dk.off_hi), dk.key);
    const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
    int i = 0;
    while (i < V - 2 && u >= cdf_row[i]) ++i;
    const int n = rot_off[i + 1] - rot_off[i];
    return rot_off[i] + (int)__umulhi(r.y, (uint32_t)n);
}

// per-slot tables shared by the sampler and the scorer: rotated class order k+1..V-1,0..k-1, prefix offsets, CDF row
struct SlotTables {
    int rot_off[CSS_CMAX + 1];
    int rot_cls[CSS_CMAX];
    float cdf[CSS_CMAX];
};

__device__ __forceinline__ void build_slot_tables(SlotTables& t, const int32_t* __restrict__ meta,
                                                  const float* __restrict__ class_cdf, int k, int V) {
    if (threadIdx.x < CSS_CMAX) {
        const int i = threadIdx.x;
        t.cdf[i] = class_cdf[k * CSS_CMAX + i];
        t.rot_cls[i] = (i < V - 1) ? meta[CSS_META_CLS_OF_SLOT + (k + 1 + i) % V] : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < V - 1; ++i) {
            t.rot_off[i] = run;
            run += meta[CSS_META_N_VALID + t.rot_cls[i]];
        }
        for (int i = V - 1; i <= CSS_CMAX; ++i) t.rot_off[i] = run;
    }
    __syncthreads();
}

// -------------------------------------------------------------------------------------------------------------------
// css_sample: materialise the draws (same device functions as the scorer uses on the fly)
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_kernel(const int32_t* __restrict__ meta, const float* __restrict__ class_cdf,
                                                     uint64_t seed, uint64_t offset, int Q, int Nn, int32_t* __restrict__ anchor_idx,
                                                     int32_t* __restrict__ neg_idx) {
    __shared__ SlotTables tb;
    const int k = blockIdx.y;
    const int V = meta[CSS_META_V];
    if (k >= V || V <= 1) return;
    const int c = meta[CSS_META_CLS_OF_SLOT + k];
    const int n_hard = meta[CSS_META_N_HARD + c];
    if (n_hard == 0) return;
    build_slot_tables(tb, meta, class_cdf, k, V);
    const DrawKey dk = make_key(seed, offset);
    const long long total = (long long)Q * (Nn + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i / (Nn + 1)), j = (int)(i - (long long)q * (Nn + 1));
        if (j == 0)
            anchor_idx[k * Q + q] = draw_anchor(dk, k, q, n_hard);
        else
            neg_idx[((size_t)k * Q + q) * Nn + (j - 1)] = draw_negative(dk, k, q, j - 1, V, tb.cdf, tb.rot_off);
    }
}

extern "C" int css_sample(const int32_t* meta, const float* class_cdf, uint64_t seed, uint64_t offset, int C, int Q, int Nn,
                          int32_t* anchor_idx, int32_t* neg_idx, void* stream) {
    CSS_CHECK_ARG(meta && class_cdf && anchor_idx && neg_idx, CSS_E_ARG, "css_sample: null pointer");
    CSS_CHECK_ARG(Q > 0 && Nn > 0 && Q < (1 << 24), CSS_E_ARG, "css_sample: bad Q/Nn");
    CSS_CHECK_ARG(C >= 1 && C <= CSS_CMAX, CSS_E_DIM, "css_sample: C must be in [1,%d]", CSS_CMAX);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(anchor_idx, 0xff, sizeof(int32_t) * (size_t)C * Q, st);
    cudaMemsetAsync(neg_idx, 0xff, sizeof(int32_t) * (size_t)C * Q * Nn, st);
    const long long total = (long long)Q * (Nn + 1);
    dim3 grid((unsigned)min((total + 255) / 256, 4096ll), C);
    sample_kernel<<<grid, 256, 0, st>>>(meta, class_cdf, seed, offset, Q, Nn, anchor_idx, neg_idx);
    CSS_CHECK_LAUNCH("css_sample", 1);
    return 0;
}

// -------------------------------------------------------------------------------------------------------------------
// css_score_ce
// -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float group_sum8(float v) {     // reduce over the 8 lanes of a group
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

__device__ __forceinline__ float dot8(const float4 (&a)[8], const float4 (&r)[8]) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        s0 = fmaf(a[i].x, r[i].x, s0);
        s0 = fmaf(a[i].y, r[i].y, s0);
        s0 = fmaf(a[i].z, r[i].z, s0);
        s0 = fmaf(a[i].w, r[i].w, s0);
        s1 = fmaf(a[i + 1].x, r[i + 1].x, s1);
        s1 = fmaf(a[i + 1].y, r[i + 1].y, s1);
        s1 = fmaf(a[i + 1].z, r[i + 1].z, s1);
        s1 = fmaf(a[i + 1].w, r[i + 1].w, s1);
    }
    return s0 + s1;
}

struct Online {            // online softmax state of one 8-lane group
    float m, l;
    float4 acc[8];
};

template <bool WANT_GRAD>
__device__ __forceinline__ void online_update(Online& st, float z, bool valid, const float4 (&r)[8]) {
    if (valid && z > st.m) {                       // group-uniform; rare after the first few rows
        const float sc = expf(st.m - z);         // m = -inf -> 0
        st.l *= sc;
        if (WANT_GRAD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                st.acc[i].x *= sc;
                st.acc[i].y *= sc;
                st.acc[i].z *= sc;
                st.acc[i].w *= sc;
            }
        }
        st.m = z;
    }
    const float wgt = valid ? expf(z - st.m) : 0.f;
    st.l += wgt;
    if (WANT_GRAD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st.acc[i].x = fmaf(wgt, r[i].x, st.acc[i].x);
            st.acc[i].y = fmaf(wgt, r[i].y, st.acc[i].y);
            st.acc[i].z = fmaf(wgt, r[i].z, st.acc[i].z);
            st.acc[i].w = fmaf(wgt, r[i].w, st.acc[i].w);
        }
    }
}

template <bool WANT_GRAD>
__global__ void __launch_bounds__(SC_THREADS) score_ce_kernel(
    const float4* __restrict__ rows_hat, const float* __restrict__ norms, const float4* __restrict__ proto_hat,
    const float* __restrict__ class_cdf, const int32_t* __restrict__ valid_list, const int32_t* __restrict__ hard_list,
    const int32_t* __restrict__ meta, const int32_t* __restrict__ anchor_idx, const int32_t* __restrict__ neg_idx, uint64_t seed,
    uint64_t offset, int N, int Q, int Nn, float temp, float* __restrict__ loss_kq, int32_t* __restrict__ anchor_px,
    float4* __restrict__ grad_anchor) {
    __shared__ SlotTables tb;
    const int k = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * SC_WARPS + warp;
    const int V = meta[CSS_META_V];
    const int c = (k < V) ? meta[CSS_META_CLS_OF_SLOT + k] : 0;
    const int n_hard = (k < V) ? meta[CSS_META_N_HARD + c] : 0;
    if (k >= V || V <= 1 || n_hard == 0) {         // absent slot / degenerate batch / no hard pixel (loss.py:116,125-130)
        if (q < Q && lane == 0) {
            loss_kq[k * Q + q] = 0.f;
            anchor_px[k * Q + q] = -1;
        }
        return;
    }
    build_slot_tables(tb, meta, class_cdf, k, V);
    if (q >= Q) return;

    const int grp = lane >> 3, l8 = lane & 7;
    const DrawKey dk = make_key(seed, offset);
    const int ai = anchor_idx ? anchor_idx[k * Q + q] : draw_anchor(dk, k, q, n_hard);
    const int pa = hard_list[(size_t)c * N + ai];
    float4 a[8];
    {
        const float4* ap = rows_hat + (size_t)pa * (CSS_D / 4) + l8;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ldg(ap + i * 8);
    }
    const float4* pp = proto_hat + (size_t)c * (CSS_D / 4) + l8;

    Online st;
    st.m = -INFINITY;
    st.l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) st.acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float z0 = 0.f, cos_pos = 0.f;

    const int n_cand = Nn + 1;                     // candidate 0 = the (updated) prototype of class c (loss.py:143-144)
    for (int base = 0; base < n_cand; base += 32) {
        // one candidate row id per lane: -1 = prototype, -2 = past the end
        int my_row = -2;
        {
            const int j = base + lane;
            if (j == 0) {
                my_row = -1;
            } else if (j < n_cand) {
                const int idx = neg_idx ? neg_idx[((size_t)k * Q + q) * Nn + (j - 1)]
                                        : draw_negative(dk, k, q, j - 1, V, tb.cdf, tb.rot_off);
                int i = 0;
                while (i < V - 2 && idx >= tb.rot_off[i + 1]) ++i;
                my_row = valid_list[(size_t)tb.rot_cls[i] * N + (idx - tb.rot_off[i])];
            }
        }
#pragma unroll 1
        for (int t = 0; t < 8; t += 2) {
            const int row0 = __shfl_sync(0xffffffffu, my_row, t * 4 + grp);
            const int row1 = __shfl_sync(0xffffffffu, my_row, (t + 1) * 4 + grp);
            const float4* p0 = (row0 >= 0) ? rows_hat + (size_t)row0 * (CSS_D / 4) + l8 : pp;
            const float4* p1 = (row1 >= 0) ? rows_hat + (size_t)row1 * (CSS_D / 4) + l8 : pp;
            float4 r0[8], r1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r0[i] = __ldg(p0 + i * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) r1[i] = __ldg(p1 + i * 8);
            const float cos0 = group_sum8(dot8(a, r0));
            const float cos1 = group_sum8(dot8(a, r1));
            const float zz0 = __fdiv_rn(cos0, temp), zz1 = __fdiv_rn(cos1, temp);
            if (row0 == -1) {
                z0 = zz0;
                cos_pos = cos0;
            }
            online_update<WANT_GRAD>(st, zz0, row0 != -2, r0);
            online_update<WANT_GRAD>(st, zz1, row1 != -2, r1);
        }
    }

    // merge the four groups (xor 8, xor 16); afterwards every lane holds the full state for its column slice
    z0 = __shfl_sync(0xffffffffu, z0, 0);
    cos_pos = __shfl_sync(0xffffffffu, cos_pos, 0);
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        const float m_o = __shfl_xor_sync(0xffffffffu, st.m, o);
        const float l_o = __shfl_xor_sync(0xffffffffu, st.l, o);
        const float M = fmaxf(st.m, m_o);
        const float s_a = (st.m == -INFINITY) ? 0.f : expf(st.m - M);
        const float s_b = (m_o == -INFINITY) ? 0.f : expf(m_o - M);
        st.l = st.l * s_a + l_o * s_b;
        if (WANT_GRAD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                st.acc[i].x = st.acc[i].x * s_a + __shfl_xor_sync(0xffffffffu, st.acc[i].x, o) * s_b;
                st.acc[i].y = st.acc[i].y * s_a + __shfl_xor_sync(0xffffffffu, st.acc[i].y, o) * s_b;
                st.acc[i].z = st.acc[i].z * s_a + __shfl_xor_sync(0xffffffffu, st.acc[i].z, o) * s_b;
                st.acc[i].w = st.acc[i].w * s_a + __shfl_xor_sync(0xffffffffu, st.acc[i].w, o) * s_b;
            }
        }
        st.m = M;
    }
    // CE with target 0: logsumexp(z) - z_0 (loss.py:147)
    if (lane == 0) {
        loss_kq[k * Q + q] = (st.m + logf(st.l)) - z0;
        anchor_px[k * Q + q] = pa;
    }
    if (WANT_GRAD) {
        // dL/da = (sum_j g_j r_hat_j - (sum_j g_j cos_j) a_hat) / max(||a||, eps),  g_j = (pi_j - [j==0]) / (Q V temp)
        // with sum_j pi_j r_hat_j = acc / l and sum_j pi_j cos_j = a_hat . (acc / l)        
        const float inv_l = 1.f / st.l;
        float4 ph[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ph[i] = __ldg(pp + i * 8);
        float sdot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st.acc[i].x *= inv_l;
            st.acc[i].y *= inv_l;
            st.acc[i].z *= inv_l;
            st.acc[i].w *= inv_l;
        }
        sdot = group_sum8(dot8(a, st.acc));
        const float scale = 1.f / ((float)Q * (float)V * temp);
        const float tt = (sdot - cos_pos) * scale;
        const float inv_na = 1.f / fmaxf(norms[pa], 1e-8f);
        if (grp == 0) {
            float4* g = grad_anchor + ((size_t)k * Q + q) * (CSS_D / 4) + l8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 o;
                o.x = ((st.acc[i].x - ph[i].x) * scale - tt * a[i].x) * inv_na;
                o.y = ((st.acc[i].y - ph[i].y) * scale - tt * a[i].y) * inv_na;
                o.z = ((st.acc[i].z - ph[i].z) * scale - tt * a[i].z) * inv_na;
                o.w = ((st.acc[i].w - ph[i].w) * scale - tt * a[i].w) * inv_na;
                g[i * 8] = o;
            }
        }
    }
}

// loss = (1/V) sum_k (1/Q) sum_q loss_kq ; exactly 0 when V <= 1 (loss.py:116-117,149).  One block, fixed-order tree.
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ loss_kq, const int32_t* __restrict__ meta, int Q,
                                                          float* __restrict__ loss) {
    __shared__ float part[8];
    const int V = meta[CSS_META_V];
    float s = 0.f;
    if (V > 1)
        for (int i = threadIdx.x; i < V * Q; i += 256) s += loss_kq[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += part[i];
        *loss = (V > 1) ? t / ((float)Q * (float)V) : 0.f;
    }
}

extern "C" int css_score_ce(const float* rows_hat, const float* norms, const float* proto_hat, const float* class_cdf,
                            const int32_t* valid_list, const int32_t* hard_list, const int32_t* meta, const int32_t* anchor_idx,
                            const int32_t* neg_idx, uint64_t seed, uint64_t offset, int N, int C, int D, int Q, int Nn, float temp,
                            float* loss_kq, int32_t* anchor_px, float* grad_anchor, float* loss, void* stream) {
    CSS_CHECK_ARG(rows_hat && norms && proto_hat && class_cdf && valid_list && hard_list && meta && loss_kq && anchor_px && loss,
                  CSS_E_ARG, "css_score_ce: null pointer");
    CSS_CHECK_ARG((anchor_idx == nullptr) == (neg_idx == nullptr), CSS_E_ARG,
                  "css_score_ce: anchor_idx and neg_idx must be fed together");
    CSS_CHECK_ARG(N > 0 && Q > 0 && Nn > 0 && Q < (1 << 24), CSS_E_ARG, "css_score_ce: bad N/Q/Nn");
    if (int e = css_check_dims(C, D)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((Q + SC_WARPS - 1) / SC_WARPS, C);
    if (grad_anchor)
        score_ce_kernel<true><<<grid, SC_THREADS, 0, st>>>((const float4*)rows_hat, norms, (const float4*)proto_hat, class_cdf,
                                                          valid_list, hard_list, meta, anchor_idx, neg_idx, seed, offset, N, Q, Nn,
                                                          temp, loss_kq, anchor_px, (float4*)grad_anchor);
    else
        score_ce_kernel<false><<<grid, SC_THREADS, 0, st>>>((const float4*)rows_hat, norms, (const float4*)proto_hat, class_cdf,
                                                           valid_list, hard_list, meta, anchor_idx, neg_idx, seed, offset, N, Q, Nn,
                                                           temp, loss_kq, anchor_px, nullptr);
    loss_reduce_kernel<<<1, 256, 0, st>>>(loss_kq, meta, Q, loss);
    CSS_CHECK_LAUNCH("css_score_ce", 2);
    return 0;
}

// -------------------------------------------------------------------------------------------------------------------
// backward: grad_rep = 0 ; grad_rep[b, :, y, x] += grad_out * grad_anchor[kq, :] at every anchor pixel
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CSS_D) grad_scatter_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ anchor_px,
                                                             const float* __restrict__ grad_anchor, int hw, float* __restrict__ grad_rep) {
    const int px = anchor_px[blockIdx.x];
    if (px < 0) return;
    const int b = px / hw, s = px - b * hw;
    const float g = __ldg(grad_out) * grad_anchor[(size_t)blockIdx.x * CSS_D + threadIdx.x];
    atomicAdd(grad_rep + ((size_t)b * CSS_D + threadIdx.x) * hw + s, g);
}

extern "C" int css_grad_scatter(const float* grad_out, const int32_t* anchor_px, const float* grad_anchor, int n_anchor, int B2,
                                int D, int h, int w, float* grad_rep, void* stream) {
    CSS_CHECK_ARG(grad_out && anchor_px && grad_anchor && grad_rep, CSS_E_ARG, "css_grad_scatter: null pointer");
    CSS_CHECK_ARG(n_anchor > 0 && B2 > 0 && h > 0 && w > 0, CSS_E_ARG, "css_grad_scatter: non-positive size");
    CSS_CHECK_ARG(D == CSS_D, CSS_E_DIM, "css_grad_scatter: D must be %d", CSS_D);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_rep, 0, sizeof(float) * (size_t)B2 * D * h * w, st);
    if (e != cudaSuccess) { css_set_error("css_grad_scatter: memset: %s", cudaGetErrorString(e)); return (int)e; }
    grad_scatter_kernel<<<n_anchor, CSS_D, 0, st>>>(grad_out, anchor_px, grad_anchor, h * w, grad_rep);
    CSS_CHECK_LAUNCH("css_grad_scatter", 1);
    return 0;
}
