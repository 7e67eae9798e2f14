// The one exchange step of the path, as a kernel over NVLink peer memory: every rank's [C, 257] block of per-class feature
// sums | counts is summed over the ranks of one node.  Replaces concat_all_gather of the representation map and the label
// map (loss.py:77,81; ddp_model.py:241-250: ~116 MB per rank) and, on one NVSwitch node, the NCCL all-reduce of the block.
//
// Every rank owns one communication buffer (cudaMalloc'ed by css_comm_alloc, exported with a CUDA IPC handle, mapped by its
// peers).  One launch, one CTA:
//   push   this rank's block is stored into slot[parity][rank] of EVERY peer's buffer (remote stores over NVLink),
//   signal after a system-scope fence, flag[parity][rank] = epoch on every peer,
//   wait   until the own buffer holds the flags of all ranks for this epoch (bounded spin: css_comm_set_timeout_ms, default
//          10 minutes like the NCCL watchdog; on expiry the block is left as the LOCAL statistics -- never poisoned -- and the
//          event is counted in a host-visible status word that css_comm_timeouts() reads without synchronising),
//   sum    the slots are added in RANK ORDER, so every rank gets bit-identical statistics (NCCL's order depends on the
//          algorithm it picks) and the prototypes cannot drift apart between ranks.
// The epoch lives in the buffer and is bumped by the kernel, so a captured CUDA graph replays correctly.  Slots alternate
// with the epoch's parity: a rank can only be one call ahead of a peer, because completing call e needs that peer's flag of
// call e, which it raises only from inside its own call e.
#include <cstring>

#include "css_common.cuh"

#define COMM_MAX_WORLD 16
#define COMM_SLOT_FLOATS (CSS_CMAX * (CSS_D + 1))
#define COMM_THREADS 1024
#define COMM_TIMEOUT_DEFAULT_MS 600000ull      // a rank may legitimately be seconds late (checkpoint write, loader respawn, GC)

struct CommHeader {
    unsigned int flags[2][COMM_MAX_WORLD];
    unsigned int epoch;
    unsigned int timeouts;                    // number of calls that gave up waiting
    unsigned int* host_status;                // pinned, mapped: [0] = timeouts, mirrored so the host can poll without a sync
    unsigned long long wait_ns;               // total time the calls spent waiting for the slowest peer (css_comm_stats)
    unsigned int pad[26];
};
static_assert(sizeof(CommHeader) == 256, "CommHeader layout");

__host__ __device__ inline size_t comm_bytes(int world) {
    return sizeof(CommHeader) + (size_t)2 * world * COMM_SLOT_FLOATS * sizeof(float);
}

__device__ __forceinline__ float* comm_slot(void* buf, int world, int parity, int src) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(buf) + sizeof(CommHeader)) + ((size_t)parity * world + src) * COMM_SLOT_FLOATS;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(COMM_THREADS) stats_allreduce_kernel(float* __restrict__ class_stats, void* local, void* const* __restrict__ peers,
                                                                       int rank, int world, int n, unsigned long long timeout_ns) {
    css_pdl_enter();
    __shared__ unsigned int s_epoch;
    __shared__ int s_ok;
    __shared__ unsigned long long s_wait;
    CommHeader* me = reinterpret_cast<CommHeader*>(local);
    if (threadIdx.x == 0) {
        s_epoch = me->epoch + 1;
        s_ok = 1;
        s_wait = 0ull;
    }
    __syncthreads();
    const unsigned int epoch = s_epoch;
    const int parity = (int)(epoch & 1u);
    // push: warp w serves peer w % world; the 32 / world warps of a peer interleave over the block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = COMM_THREADS / 32;
    {
        const int p = warp % world, sub = warp / world, share = n_warps / world;      // `share` warps split the block for one peer
        if (sub < share) {
            float* dst = comm_slot(peers[p], world, parity, rank);
            for (int i = sub * 32 + lane; i < n; i += share * 32) dst[i] = class_stats[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        CommHeader* peer = reinterpret_cast<CommHeader*>(peers[threadIdx.x]);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&peer->flags[parity][rank]), "r"(epoch) : "memory");
        // wait for rank `threadIdx.x`'s block to have landed here
        const unsigned int* f = &me->flags[parity][threadIdx.x];
        const unsigned long long t0 = global_ns();
        unsigned int v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v == epoch) break;
            if (global_ns() - t0 > timeout_ns) {
                s_ok = 0;
                break;
            }
        } while (true);
        atomicMax(&s_wait, global_ns() - t0);
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    if (ok) {
        for (int i = threadIdx.x; i < n; i += COMM_THREADS) {
            float s = 0.f;
            for (int p = 0; p < world; ++p) s += __ldcg(comm_slot(local, world, parity, p) + i);     // L2: the slots were written remotely
            class_stats[i] = s;
        }
    }   // else: a peer never arrived -- class_stats keeps this rank's own statistics (a valid, rank-local prototype update)
    if (threadIdx.x == 0) {
        me->epoch = epoch;
        me->wait_ns += s_wait;
        if (!ok) {
            me->timeouts += 1;
            if (me->host_status) {
                *reinterpret_cast<volatile unsigned int*>(me->host_status) = me->timeouts;
                __threadfence_system();
            }
        }
    }
}

extern "C" size_t css_comm_bytes(int world) { return (world >= 1 && world <= COMM_MAX_WORLD) ? comm_bytes(world) : 0; }

// host-side registry of the buffers this process allocated: device buffer -> its pinned status word
#define COMM_MAX_LOCAL 64
static struct { void* dev; unsigned int* host; } g_comm[COMM_MAX_LOCAL];
static unsigned long long g_timeout_ms = COMM_TIMEOUT_DEFAULT_MS;

static int comm_find(void* buffer) {
    for (int i = 0; i < COMM_MAX_LOCAL; ++i)
        if (g_comm[i].dev == buffer) return i;
    return -1;
}

extern "C" int css_comm_set_timeout_ms(unsigned long long ms) {
    CSS_CHECK_ARG(ms > 0, CSS_E_ARG, "css_comm_set_timeout_ms: timeout must be positive");
    g_timeout_ms = ms;
    return 0;
}

extern "C" int css_comm_alloc(int world, void** buffer) {
    CSS_CHECK_ARG(buffer && world >= 1 && world <= COMM_MAX_WORLD, CSS_E_ARG, "css_comm_alloc: world must be in [1,%d]", COMM_MAX_WORLD);
    const int slot = comm_find(nullptr);
    CSS_CHECK_ARG(slot >= 0, CSS_E_SIZE, "css_comm_alloc: more than %d live communication buffers", COMM_MAX_LOCAL);
    void* p = nullptr;
    unsigned int* hs = nullptr;
    unsigned int* hs_dev = nullptr;
    cudaError_t e = cudaMalloc(&p, comm_bytes(world));
    if (e == cudaSuccess) e = cudaMemset(p, 0, comm_bytes(world));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&hs, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) {
        memset(hs, 0, 64);
        e = cudaHostGetDevicePointer((void**)&hs_dev, hs, 0);
    }
    if (e == cudaSuccess) e = cudaMemcpy(&reinterpret_cast<CommHeader*>(p)->host_status, &hs_dev, sizeof(hs_dev), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        if (hs) cudaFreeHost(hs);
        css_set_error("css_comm_alloc: %s", cudaGetErrorString(e));
        return (int)e;
    }
    g_comm[slot].dev = p;
    g_comm[slot].host = hs;
    *buffer = p;
    return 0;
}

extern "C" int css_comm_free(void* buffer) {
    if (!buffer) return 0;
    const int slot = comm_find(buffer);
    cudaError_t e = cudaFree(buffer);
    if (slot >= 0) {
        if (g_comm[slot].host) cudaFreeHost(g_comm[slot].host);
        g_comm[slot].dev = nullptr;
        g_comm[slot].host = nullptr;
    }
    if (e != cudaSuccess) { css_set_error("css_comm_free: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// diagnostics of a local buffer: out[0] = completed calls (epoch), out[1] = timeouts, out[2] = total ns spent waiting for peers.
// Reads device memory with a blocking copy: never call it inside a timed region.
extern "C" int css_comm_stats(void* buffer, unsigned long long* out3_host) {
    CSS_CHECK_ARG(buffer && out3_host && comm_find(buffer) >= 0, CSS_E_ARG, "css_comm_stats: not a buffer of css_comm_alloc");
    CommHeader h;
    cudaError_t e = cudaMemcpy(&h, buffer, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { css_set_error("css_comm_stats: %s", cudaGetErrorString(e)); return (int)e; }
    out3_host[0] = h.epoch;
    out3_host[1] = h.timeouts;
    out3_host[2] = h.wait_ns;
    return 0;
}

// number of css_stats_allreduce calls on `buffer` that gave up waiting for a peer; reads pinned host memory, never synchronises
extern "C" int css_comm_timeouts(void* buffer) {
    const int slot = buffer ? comm_find(buffer) : -1;
    CSS_CHECK_ARG(slot >= 0, CSS_E_ARG, "css_comm_timeouts: not a buffer of css_comm_alloc");
    return (int)*reinterpret_cast<volatile unsigned int*>(g_comm[slot].host);
}

extern "C" int css_comm_export(void* buffer, unsigned char* handle64) {
    CSS_CHECK_ARG(buffer && handle64, CSS_E_ARG, "css_comm_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, buffer);
    if (e != cudaSuccess) { css_set_error("css_comm_export: %s", cudaGetErrorString(e)); return (int)e; }
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int css_comm_open(const unsigned char* handle64, void** peer_buffer) {
    CSS_CHECK_ARG(handle64 && peer_buffer, CSS_E_ARG, "css_comm_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { (void)cudaGetLastError(); css_set_error("css_comm_open: %s", cudaGetErrorString(e)); return (int)e; }
    *peer_buffer = p;
    return 0;
}

extern "C" int css_comm_close(void* peer_buffer) {
    if (!peer_buffer) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(peer_buffer);
    if (e != cudaSuccess) { css_set_error("css_comm_close: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int css_stats_allreduce(float* class_stats, void* local_buffer, void* const* peer_buffers, int rank, int world, int C, int D,
                                   void* stream) {
    CSS_CHECK_ARG(class_stats && local_buffer && peer_buffers, CSS_E_ARG, "css_stats_allreduce: null pointer");
    CSS_CHECK_ARG(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, CSS_E_ARG, "css_stats_allreduce: bad rank %d / world %d",
                  rank, world);
    if (int e = css_check_dims(C, D)) return e;
    css_launch(stats_allreduce_kernel, dim3(1), dim3(COMM_THREADS), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), class_stats, local_buffer, peer_buffers, rank, world, C * (D + 1),
                                                                         g_timeout_ms * 1000000ull);
    CSS_CHECK_LAUNCH("css_stats_allreduce", 1);
    return 0;
}
