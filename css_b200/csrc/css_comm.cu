// The one exchange step of the path, as a kernel over NVLink peer memory: every rank's [C, 257] block of per-class feature
// sums | counts is summed over the ranks of one node.  Replaces concat_all_gather of the representation map and the label
// map (loss.py:77,81; ddp_model.py:241-250: ~116 MB per rank) and, on one NVSwitch node, the NCCL all-reduce of the block.
//
// Every rank owns one communication buffer (cudaMalloc'ed by css_comm_alloc, exported with a CUDA IPC handle, mapped by its
// peers).  One launch of `world` CTAs; CTA j serves peer (rank + j) % world and slice j of the block:
//   push   the block is read into registers (one L2 round trip) and stored into slot[parity][rank] of the CTA's peer (remote
//          stores over NVLink, one SM per peer so the stores of all peers stream out concurrently),
//   signal after a system-scope fence, flag[parity][rank] = epoch on that peer,
//   wait   until the own buffer holds the flags of all ranks for this epoch and every local CTA has read the block (bounded spin:
//          css_comm_set_timeout_ms, default 10 minutes like the NCCL watchdog; on expiry nothing is poisoned -- a CTA that gives up
//          leaves its slice of class_stats as the LOCAL statistics -- and the event is counted in a host-visible status word that
//          css_comm_timeouts() reads without synchronising),
//   sum    the CTA adds its slice of the slots in RANK ORDER, so every rank gets bit-identical statistics (NCCL's order depends
//          on the algorithm it picks) and the prototypes cannot drift apart between ranks.
// Measured on 8 B200 (V321): the first version (one CTA, a dependent L2 round trip per element in the push
// and the sum loops) cost ~30 us per step, one CTA with the loads batched 18 us (7.5 us of it issuing the stores to seven peers
// from one SM); an LL-style version ({value, epoch} pairs, no fence, polling every pair) 33 us.
// The epoch lives in the buffer and is bumped by the last CTA of the launch, so a captured CUDA graph replays correctly.  Slots
// alternate with the epoch's parity: a rank can only be one call ahead of a peer, because completing call e needs that peer's flag
// of call e, which it raises only from inside its own call e.
#include <cstring>

#include "css_common.cuh"

#define COMM_MAX_WORLD 16
#define COMM_SLOT_FLOATS (CSS_CMAX * (CSS_D + 1))
#define COMM_THREADS 1024
#define COMM_PER_THREAD ((COMM_SLOT_FLOATS + COMM_THREADS - 1) / COMM_THREADS + ((COMM_SLOT_FLOATS + COMM_THREADS - 1) / COMM_THREADS) % 2)   // even
#define COMM_TIMEOUT_DEFAULT_MS 600000ull      // a rank may legitimately be seconds late (checkpoint write, loader respawn, GC)

struct CommHeader {
    unsigned int flags[2][COMM_MAX_WORLD];
    unsigned int epoch;
    unsigned int timeouts;                    // number of calls that gave up waiting
    unsigned int* host_status;                // pinned, mapped: [0] = timeouts, mirrored so the host can poll without a sync
    unsigned long long wait_ns;               // total time the calls spent waiting for the slowest peer (css_comm_stats)
    unsigned long long push_ns;               // total time from kernel entry until the pushed block was fenced (remote stores acknowledged)
    unsigned long long total_ns;              // total time inside the kernel
    unsigned int loaded;                      // local CTAs that have read class_stats in this call (the sums overwrite it in place)
    unsigned int ticket;                      // local CTAs that have finished this call: the last one bumps the epoch, resets both
    unsigned int pad[20];
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

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int WMAX>
__global__ void __launch_bounds__(COMM_THREADS) stats_allreduce_kernel(float* __restrict__ class_stats, void* local, void* const* __restrict__ peers,
                                                                       int rank, int world, int n, unsigned long long timeout_ns) {
    css_pdl_enter();
    __shared__ int s_ok;
    __shared__ unsigned long long s_wait;
    CommHeader* me = reinterpret_cast<CommHeader*>(local);
    const unsigned long long t_enter = global_ns();
    const int cta = blockIdx.x, peer_id = (rank + cta) % world;        // CTA 0 fills the own buffer's slot
    // the epoch, the peer's address and the block are independent reads: all in flight together (one L2 round trip, not three).
    // Every thread reads the epoch itself (a broadcast load); it is bumped only after every CTA of this launch has finished.
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(&me->epoch) + 1u;
    void* const peer_buf = peers[peer_id];
    float own[COMM_PER_THREAD];            // thread t owns elements t, t + 1024, ...
#pragma unroll
    for (int j = 0; j < COMM_PER_THREAD; ++j) {
        const int i = threadIdx.x + j * COMM_THREADS;
        own[j] = (i < n) ? class_stats[i] : 0.f;
    }
    if (threadIdx.x == 0) {
        s_ok = 1;
        s_wait = 0ull;
    }
    const int parity = (int)(epoch & 1u);
    {
        float* dst = comm_slot(peer_buf, world, parity, rank);
#pragma unroll
        for (int j = 0; j < COMM_PER_THREAD; ++j) {
            const int i = threadIdx.x + j * COMM_THREADS;
            if (i < n) dst[i] = own[j];
        }
    }
    __syncthreads();                      // every thread's loads have returned and its stores are issued
    unsigned long long t_pushed = 0ull;
    if (threadIdx.x == 0) {
        atomicAdd(&me->loaded, 1u);       // this CTA no longer needs class_stats
        // ONE system-scope fence per CTA: it is cumulative over the barrier above, so the whole CTA's stores are ordered before
        // the flag (1024 fences, one per thread, cost several microseconds and add nothing)
        __threadfence_system();
        CommHeader* peer = reinterpret_cast<CommHeader*>(peer_buf);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&peer->flags[parity][rank]), "r"(epoch) : "memory");
        t_pushed = global_ns();
    }
    // wait: thread r < world for rank r's block to have landed here, thread `world` for the local CTAs to have read class_stats
    if (threadIdx.x <= world) {
        const unsigned int* f = (threadIdx.x < world) ? &me->flags[parity][threadIdx.x] : &me->loaded;
        const unsigned int want = (threadIdx.x < world) ? epoch : (unsigned int)world;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(f) != want) {
            if (global_ns() - t0 > timeout_ns) {
                s_ok = 0;
                break;
            }
        }
        atomicMax(&s_wait, global_ns() - t0);
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    if (ok) {
        // this CTA's slice, summed in rank order; all ranks' values of an element are in flight together (L2: written remotely)
        const int chunk = (n + world - 1) / world, lo = cta * chunk, hi = min(n, lo + chunk);
        const float* slot0 = comm_slot(local, world, parity, 0);
        for (int i = lo + threadIdx.x; i < hi; i += COMM_THREADS) {
            float v[WMAX];
#pragma unroll
            for (int p = 0; p < WMAX; ++p) v[p] = (p < world) ? __ldcg(slot0 + (size_t)p * COMM_SLOT_FLOATS + i) : 0.f;
            float s = 0.f;
#pragma unroll
            for (int p = 0; p < WMAX; ++p)
                if (p < world) s += v[p];
            class_stats[i] = s;
        }
    }   // else: a peer never arrived -- this slice of class_stats keeps the rank's own statistics (valid, rank-local)
    __syncthreads();
    if (threadIdx.x == 0) {
        if (cta == 0) {                   // diagnostics from the CTA that serves the own buffer
            me->wait_ns += s_wait;
            me->push_ns += t_pushed - t_enter;
            me->total_ns += global_ns() - t_enter;
        }
        if (!ok) atomicAdd(&me->timeouts, 0x10000u);          // high half: CTAs that gave up in this call
        __threadfence();
        if (atomicAdd(&me->ticket, 1u) == (unsigned int)world - 1u) {      // last CTA of the launch
            __threadfence();
            const unsigned int t = *reinterpret_cast<volatile unsigned int*>(&me->timeouts);
            if (t >> 16) {                 // one timed-out CALL, however many CTAs gave up
                me->timeouts = (t & 0xffffu) + 1u;
                if (me->host_status) {
                    *reinterpret_cast<volatile unsigned int*>(me->host_status) = me->timeouts;
                    __threadfence_system();
                }
            }
            me->loaded = 0u;
            me->ticket = 0u;
            me->epoch = epoch;
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

// diagnostics of a local buffer: out[0] = completed calls (epoch), out[1] = timeouts, out[2] = total ns spent waiting for peers,
// out[3] = total ns from kernel entry to the fenced push, out[4] = total ns inside the kernel.
// Reads device memory with a blocking copy: never call it inside a timed region.
extern "C" int css_comm_stats(void* buffer, unsigned long long* out5_host) {
    unsigned long long* out3_host = out5_host;
    CSS_CHECK_ARG(buffer && out3_host && comm_find(buffer) >= 0, CSS_E_ARG, "css_comm_stats: not a buffer of css_comm_alloc");
    CommHeader h;
    cudaError_t e = cudaMemcpy(&h, buffer, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { css_set_error("css_comm_stats: %s", cudaGetErrorString(e)); return (int)e; }
    out3_host[0] = h.epoch;
    out3_host[1] = h.timeouts;
    out3_host[2] = h.wait_ns;
    out5_host[3] = h.push_ns;
    out5_host[4] = h.total_ns;
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
    const int n = C * (D + 1);
    const unsigned long long t_ns = g_timeout_ms * 1000000ull;
    cudaStream_t st = (cudaStream_t)stream;
    if (world <= 2) css_launch(stats_allreduce_kernel<2>, dim3(world), dim3(COMM_THREADS), (size_t)0, st, class_stats, local_buffer, peer_buffers, rank, world, n, t_ns);
    else if (world <= 4) css_launch(stats_allreduce_kernel<4>, dim3(world), dim3(COMM_THREADS), (size_t)0, st, class_stats, local_buffer, peer_buffers, rank, world, n, t_ns);
    else if (world <= 8) css_launch(stats_allreduce_kernel<8>, dim3(world), dim3(COMM_THREADS), (size_t)0, st, class_stats, local_buffer, peer_buffers, rank, world, n, t_ns);
    else css_launch(stats_allreduce_kernel<COMM_MAX_WORLD>, dim3(world), dim3(COMM_THREADS), (size_t)0, st, class_stats, local_buffer, peer_buffers, rank, world, n, t_ns);
    CSS_CHECK_LAUNCH("css_stats_allreduce", 1);
    return 0;
}
