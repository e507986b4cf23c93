// Data-parallel gradient exchange over NVLink peer memory.
//
// The reference trains on one device (core/model.py:148-171: one GradientTape, one apply_gradients over the 48
// variables); data parallel, every rank holds the same flat buffer [coarse gradient | fine gradient | loss, 0, 0, 0]
// and the step needs its SUM over the ranks before the replicated Adam. Instead of handing that buffer to NCCL, the
// ranks map each other's buffer (CUDA IPC over NVLink / NVSwitch) and ONE kernel per rank does the whole exchange:
//
//   barrier A   "my local gradient is complete" -> a flag in every peer's block; wait for every peer's flag
//   reduce      rank r owns slice r of the buffer: it loads that slice from every rank over NVLink (16-byte loads, all
//               ranks' loads of an element in flight together), sums them in rank order 0..W-1
//   broadcast   ... and stores the sum into slice r of EVERY rank's buffer, in place: slice r of any buffer is read
//               and written by rank r only, and an element is written after it was read by the same thread.
//               With an NVSwitch multicast mapping of the block (NVLS) both steps are done by the switch:
//               multimem.ld_reduce returns the sum over all GPUs, multimem.st replicates it into all of them
//   barrier B   "my slice has landed everywhere": the CTAs of this GPU are ordered by a gpu-scope fence and a counter,
//               the LAST one releases a flag to every peer at system scope (cumulative over that chain) and waits for
//               every peer's flag, so the kernel (and with it the stream) completes only when the local buffer holds
//               the full sum
//
// Every element is summed by exactly one rank (or once by the switch) and replicated, so all replicas receive
// bit-identical sums (they stay bit-identical replicas) and the result does not depend on timing. Flags carry a
// monotonically increasing epoch that lives in device memory, so a launch captured in a CUDA graph replays correctly.
// 4.77 MB over 8 GPUs is latency, not bandwidth: 25 us with NVLS, 28 us unicast, against 54 us for an NCCL all-reduce
// (DESIGN.md 4.2c has the phase timings the kernel stamps into its header).
//
// Optional epilogue (adam != 0): instead of leaving after its slice, every CTA waits for barrier B and then applies
// the fused Adam step (optim.cu) to its share of the parameters -- the exchange and the optimizer in one launch.
#include "common.cuh"

#include <math.h>

namespace nb {

constexpr int kPeerMaxWorld = 8;        // one NVSwitch box
constexpr int kPeerHeaderBytes = 4096;
constexpr int kPeerThreads = 256;

struct PeerHeader {                       // at offset 0 of every rank's block; the exchanged floats start at kPeerHeaderBytes
    uint32_t arrive[2][kPeerMaxWorld];    // [barrier][source rank]: epoch of that rank's latest arrival
    uint32_t epoch;                       // exchanges completed by this rank
    uint32_t done_ctas;                   // CTAs of the running exchange that have finished their slice
    uint32_t status;                      // != 0: a wait timed out (the kernel traps right after setting it)
    uint32_t pad;
    unsigned long long stamp[7];          // %globaltimer of the latest exchange (nerfb200_peer_profile): start, barrier A
};                                        // passed, own slice done [CTA 0]; all CTAs done, barrier B passed, end [last CTA];
                                          // [6]: CTA 0's loads have returned and its stores are issued (before the fence)
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "header");

struct AdamArgs {
    float* p; float* m; float* v;
    const int64_t* step_dev;              // device-resident iteration counter, or NULL
    float lr_t;
    int64_t n;
};

struct PeerParams {
    uint8_t* base[kPeerMaxWorld];         // every rank's block; base[rank] is local memory
    uint8_t* mc;                          // the same block through an NVSwitch multicast mapping (NVLS), or NULL
    int world, rank;
    int64_t n4;                           // 16-byte units exchanged
    unsigned long long timeout_ns;
    int adam;
    AdamArgs ad;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {     // never served from a stale cache line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// NVLS: one load returns the SUM of the addressed 16 bytes over every GPU of the multicast group (the switch reduces),
// one store writes them into every GPU's copy (the switch replicates)
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f4(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// threads [0, world) of the CTA each wait for one source rank's flag; everybody leaves together
__device__ __forceinline__ void wait_flags(PeerHeader* me, int barrier, int world, uint32_t epoch, unsigned long long timeout_ns) {
    if ((int)threadIdx.x < world) {
        const uint32_t* f = &me->arrive[barrier][threadIdx.x];
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
                me->status = 1u + (uint32_t)barrier;
                __threadfence_system();
                __trap();          // a peer never arrived (crashed, or the ranks disagree about the step): fail loudly
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const PeerParams P) {
    PeerHeader* me = reinterpret_cast<PeerHeader*>(P.base[P.rank]);
    __shared__ uint32_t s_epoch;
    __shared__ int s_last;
    __shared__ float s_lr_t;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(&me->epoch) + 1u;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const bool stamper = blockIdx.x == 0 && threadIdx.x == 0;
    if (stamper) me->stamp[0] = global_ns();

    // ---- barrier A. The local gradient was written by earlier kernels of this stream: complete and visible in this
    // GPU's memory (where the peers' loads are served) when this kernel starts.
    if (blockIdx.x == 0 && (int)threadIdx.x < P.world)
        st_release_sys(&reinterpret_cast<PeerHeader*>(P.base[threadIdx.x])->arrive[0][P.rank], epoch);
    wait_flags(me, 0, P.world, epoch, P.timeout_ns);
    if (stamper) me->stamp[1] = global_ns();

    // ---- reduce + broadcast this rank's slice
    const int64_t per = (P.n4 + P.world - 1) / P.world;
    const int64_t lo = per * P.rank, hi = (lo + per < P.n4) ? lo + per : P.n4;
    if (P.mc) {          // NVLS: the switch sums the slice and replicates the result (the order of the sum is the switch's)
        float4* mc = reinterpret_cast<float4*>(P.mc + kPeerHeaderBytes);
        for (int64_t i = lo + (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < hi; i += (int64_t)gridDim.x * kPeerThreads)
            multimem_st_f4(mc + i, multimem_ld_reduce_f4(mc + i));
    } else {
        for (int64_t i = lo + (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < hi; i += (int64_t)gridDim.x * kPeerThreads) {
            float4 v[kPeerMaxWorld];
#pragma unroll
            for (int p = 0; p < kPeerMaxWorld; ++p)
                if (p < P.world) v[p] = ld_sys_f4(reinterpret_cast<const float4*>(P.base[p] + kPeerHeaderBytes) + i);
            float4 s = v[0];
#pragma unroll
            for (int p = 1; p < kPeerMaxWorld; ++p)
                if (p < P.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
#pragma unroll
            for (int p = 0; p < kPeerMaxWorld; ++p)
                if (p < P.world) st_sys_f4(reinterpret_cast<float4*>(P.base[p] + kPeerHeaderBytes) + i, s);
        }
    }
    if (stamper) me->stamp[6] = global_ns();
    // Only the CTAs that had a part of the slice are counted (with the Adam epilogue the grid is four times wider than
    // the slice needs; 592 fences and atomics instead of 146 cost 5 us).
    int64_t wk = (hi - lo + kPeerThreads - 1) / kPeerThreads;
    const unsigned working = (unsigned)(wk < 1 ? 1 : (wk > (int64_t)gridDim.x ? (int64_t)gridDim.x : wk));
    __syncthreads();                 // the CTA's peer stores happen before thread 0's fence (cumulativity over the CTA
    if (threadIdx.x == 0) {          // barrier), which is before the CTA is counted as done
        s_last = 0;
        if (blockIdx.x < working) {
            // gpu scope: this fence and the counter order the CTAs of THIS GPU; the one system-scope release is the last
            // CTA's flag store below, and it is cumulative over this chain (a system fence here, per CTA, took ~10 us
            // on its own: N=2 exchange 23.4 -> 18.3 us, profiles/r2ab_exchange_fence_scope_n2.json)
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            if (stamper) me->stamp[2] = global_ns();
            s_last = (atomicAdd(&me->done_ctas, 1u) == working - 1) ? 1 : 0;
            asm volatile("fence.acq_rel.gpu;" ::: "memory");       // the last CTA has seen every other CTA's stores
        }
    }
    __syncthreads();

    // ---- barrier B: the LAST CTA to finish tells every peer that this rank's slice has landed
    if (s_last && threadIdx.x == 0) me->stamp[3] = global_ns();
    if (s_last && (int)threadIdx.x < P.world)
        st_release_sys(&reinterpret_cast<PeerHeader*>(P.base[threadIdx.x])->arrive[1][P.rank], epoch);
    if (!P.adam) {
        if (!s_last) return;
        wait_flags(me, 1, P.world, epoch, P.timeout_ns);      // the kernel ends when every peer's slice is here
        if (threadIdx.x == 0) { me->stamp[4] = me->stamp[5] = global_ns(); me->done_ctas = 0u; me->epoch = epoch; }
        return;
    }
    // ---- fused Adam epilogue: every CTA needs the full sum (grid <= SM count: all CTAs are resident, none waits for a
    // CTA that cannot run)
    wait_flags(me, 1, P.world, epoch, P.timeout_ns);
    if (s_last && threadIdx.x == 0) me->stamp[4] = global_ns();
    if (P.ad.step_dev) {      // device-resident iteration counter (CUDA-graph replays)
        if (threadIdx.x == 0) s_lr_t = adam_lr_t(P.ad.step_dev[0]);
        __syncthreads();
    }
    const float lr_t = P.ad.step_dev ? s_lr_t : P.ad.lr_t;
    const float4* g4 = reinterpret_cast<const float4*>(P.base[P.rank] + kPeerHeaderBytes);
    float4* p4 = reinterpret_cast<float4*>(P.ad.p);
    float4* m4 = reinterpret_cast<float4*>(P.ad.m);
    float4* v4 = reinterpret_cast<float4*>(P.ad.v);
    for (int64_t i = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < P.ad.n / 4; i += (int64_t)gridDim.x * kPeerThreads) {
        const float4 g = ld_sys_f4(g4 + i);
        float4 m = m4[i], v = v4[i], p = p4[i];
        adam_update(p.x, m.x, v.x, g.x, lr_t);
        adam_update(p.y, m.y, v.y, g.y, lr_t);
        adam_update(p.z, m.z, v.z, g.z, lr_t);
        adam_update(p.w, m.w, v.w, g.w, lr_t);
        m4[i] = m; v4[i] = v; p4[i] = p;
    }
    if (s_last && threadIdx.x == 0) { me->stamp[5] = global_ns(); me->done_ctas = 0u; me->epoch = epoch; }
}

}  // namespace nb

using namespace nb;

struct nerfb200_peer {
    int world = 0, rank = 0, device = 0, connected = 0;
    int timeout_s = 120;              // how long a rank waits for its peers inside the kernel before it traps
    int owned = 1;                    // the blocks were allocated / mapped by this library (CUDA IPC)
    int occupancy = 0;                // resident CTAs of the exchange kernel per SM (queried at the first launch)
    int64_t n = 0;
    uint8_t* base[kPeerMaxWorld] = {};
    uint8_t* mc = nullptr;
};

extern "C" {

int nerfb200_peer_create(int world, int rank, int64_t n_floats, nerfb200_peer** peer) {
    NB_CHECK_ARG(peer != nullptr, "peer_create: NULL output");
    *peer = nullptr;
    NB_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer_create: world must be in [1,%d] and rank in [0,world)", kPeerMaxWorld);
    NB_CHECK_ARG(n_floats > 0 && n_floats % 4 == 0, "peer_create: the exchanged buffer must be a positive multiple of 4 floats");
    nerfb200_peer* p = new nerfb200_peer();
    p->world = world; p->rank = rank; p->n = n_floats;
    cudaError_t e = cudaGetDevice(&p->device);
    void* mem = nullptr;
    const size_t bytes = (size_t)kPeerHeaderBytes + (size_t)n_floats * 4;
    if (e == cudaSuccess) e = cudaMalloc(&mem, bytes);
    if (e == cudaSuccess) e = cudaMemset(mem, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();     // zeroed before any peer can see the handle
    if (e != cudaSuccess) {
        set_error("peer_create: %s", cudaGetErrorString(e));
        if (mem) cudaFree(mem);
        delete p;
        return (int)e;
    }
    p->base[rank] = (uint8_t*)mem;
    *peer = p;
    return 0;
}

int nerfb200_peer_attach(int world, int rank, int64_t n_floats, void* const* blocks, void* multicast_block, nerfb200_peer** peer) {
    NB_CHECK_ARG(peer != nullptr, "peer_attach: NULL output");
    *peer = nullptr;
    NB_CHECK_ARG(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer_attach: world must be in [1,%d] and rank in [0,world)", kPeerMaxWorld);
    NB_CHECK_ARG(n_floats > 0 && n_floats % 4 == 0 && blocks, "peer_attach: the exchanged buffer must be a positive multiple of 4 floats");
    for (int r = 0; r < world; ++r)
        NB_CHECK_ARG(blocks[r] && ((uintptr_t)blocks[r] & 15) == 0, "peer_attach: block %d is NULL or not 16-byte aligned", r);
    NB_CHECK_ARG(((uintptr_t)multicast_block & 15) == 0, "peer_attach: multicast block not 16-byte aligned");
    nerfb200_peer* p = new nerfb200_peer();
    p->world = world; p->rank = rank; p->n = n_floats; p->owned = 0; p->connected = 1;
    p->mc = (uint8_t*)multicast_block;
    for (int r = 0; r < world; ++r) p->base[r] = (uint8_t*)blocks[r];
    cudaError_t e = cudaGetDevice(&p->device);
    if (e != cudaSuccess) {
        set_error("peer_attach: %s", cudaGetErrorString(e));
        delete p;
        return (int)e;
    }
    *peer = p;
    return 0;
}

int nerfb200_peer_buffer(nerfb200_peer* peer, float** buffer) {
    NB_CHECK_ARG(peer && buffer, "peer_buffer: NULL argument");
    *buffer = reinterpret_cast<float*>(peer->base[peer->rank] + kPeerHeaderBytes);
    return 0;
}

int nerfb200_peer_handle(nerfb200_peer* peer, unsigned char* handle64) {
    NB_CHECK_ARG(peer && handle64 && peer->owned, "peer_handle: NULL argument, or a handle attached to the caller's memory");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    NB_CUDA(cudaIpcGetMemHandle(&h, peer->base[peer->rank]));
    memcpy(handle64, &h, 64);
    return 0;
}

int nerfb200_peer_connect(nerfb200_peer* peer, const unsigned char* handles) {
    NB_CHECK_ARG(peer && handles, "peer_connect: NULL argument");
    NB_CHECK_ARG(!peer->connected && peer->owned, "peer_connect: already connected");
    int dev = -1;
    NB_CUDA(cudaGetDevice(&dev));
    NB_CHECK_ARG(dev == peer->device, "peer_connect: created on device %d, current device is %d", peer->device, dev);
    for (int r = 0; r < peer->world; ++r) {
        if (r == peer->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* mem = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&mem, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("peer_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (q != peer->rank && peer->base[q]) { cudaIpcCloseMemHandle(peer->base[q]); peer->base[q] = nullptr; }
            return (int)e;
        }
        peer->base[r] = (uint8_t*)mem;
    }
    peer->connected = 1;
    return 0;
}

static int peer_launch(nerfb200_peer* peer, const AdamArgs* ad, void* stream) {
    NB_CHECK_ARG(peer != nullptr, "peer_allreduce: NULL handle");
    if (!peer->connected && peer->world > 1) {
        set_error("peer_allreduce: peer_connect has not been called");
        return NERFB200_ESTATE;
    }
    int dev = -1;
    NB_CUDA(cudaGetDevice(&dev));
    NB_CHECK_ARG(dev == peer->device, "peer_allreduce: created on device %d, current device is %d", peer->device, dev);
    PeerParams P = {};
    for (int r = 0; r < peer->world; ++r) P.base[r] = peer->base[r];
    P.mc = peer->mc;
    P.world = peer->world; P.rank = peer->rank; P.n4 = peer->n / 4;
    P.timeout_ns = (unsigned long long)peer->timeout_s * 1000000000ull;
    P.adam = ad ? 1 : 0;
    if (ad) P.ad = *ad;
    const int64_t per = (P.n4 + P.world - 1) / P.world;
    const int64_t units = (ad && ad->n / 4 > per) ? ad->n / 4 : per;       // 16-byte units of the widest phase
    // every CTA of the grid must be resident at once: with the Adam epilogue all of them wait for the peers' flags
    if (peer->occupancy <= 0) {         // once per handle
        int occ = 1;
        NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, peer_allreduce_kernel, kPeerThreads, 0));
        peer->occupancy = occ > 4 ? 4 : (occ < 1 ? 1 : occ);
    }
    const int occ = peer->occupancy;
    int64_t grid = (units + kPeerThreads - 1) / kPeerThreads;
    if (grid > (int64_t)num_sms() * occ) grid = (int64_t)num_sms() * occ;
    if (grid < 1) grid = 1;
    peer_allreduce_kernel<<<(unsigned)grid, kPeerThreads, 0, (cudaStream_t)stream>>>(P);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_peer_allreduce(nerfb200_peer* peer, void* stream) { return peer_launch(peer, nullptr, stream); }

int nerfb200_peer_allreduce_adam(nerfb200_peer* peer, int64_t n, float* params, float* m, float* v, int64_t iterations,
                                 const int64_t* step_state, void* stream) {
    NB_CHECK_ARG(peer && params && m && v && iterations >= 0, "peer_allreduce_adam: bad arguments");
    NB_CHECK_ARG(n > 0 && n <= peer->n && n % 4 == 0, "peer_allreduce_adam: n must be a multiple of 4 in (0, exchanged floats]");
    NB_CHECK_ARG((((uintptr_t)params | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "peer_allreduce_adam: params, m and v must be 16-byte aligned");
    AdamArgs ad;
    ad.p = params; ad.m = m; ad.v = v; ad.step_dev = step_state; ad.n = n;
    ad.lr_t = adam_lr_t(iterations);
    return peer_launch(peer, &ad, stream);
}

int nerfb200_peer_set_timeout(nerfb200_peer* peer, int seconds) {
    NB_CHECK_ARG(peer && seconds >= 1, "peer_set_timeout: bad arguments");
    peer->timeout_s = seconds;
    return 0;
}

int nerfb200_peer_status(nerfb200_peer* peer, int* status) {
    NB_CHECK_ARG(peer && status, "peer_status: NULL argument");
    uint32_t s = 0;
    NB_CUDA(cudaMemcpy(&s, peer->base[peer->rank] + offsetof(PeerHeader, status), 4, cudaMemcpyDeviceToHost));
    *status = (int)s;
    return 0;
}

int nerfb200_peer_profile(nerfb200_peer* peer, unsigned long long* ns7) {
    NB_CHECK_ARG(peer && ns7, "peer_profile: NULL argument");
    NB_CUDA(cudaMemcpy(ns7, peer->base[peer->rank] + offsetof(PeerHeader, stamp), 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

int nerfb200_peer_disconnect(nerfb200_peer* peer) {
    NB_CHECK_ARG(peer != nullptr, "peer_disconnect: NULL handle");
    for (int r = 0; r < peer->world && peer->owned; ++r) {
        if (r == peer->rank || !peer->base[r]) continue;
        cudaIpcCloseMemHandle(peer->base[r]);
        peer->base[r] = nullptr;
    }
    peer->connected = 0;
    return 0;
}

int nerfb200_peer_destroy(nerfb200_peer* peer) {
    if (!peer) return 0;
    nerfb200_peer_disconnect(peer);
    if (peer->owned && peer->base[peer->rank]) cudaFree(peer->base[peer->rank]);
    delete peer;
    return 0;
}

}  // extern "C"
