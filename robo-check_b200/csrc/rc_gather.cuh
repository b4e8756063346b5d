// rc_gather.cuh -- all-gather of the validity bitmask through peer-mapped memory (NVLink / NVSwitch),
// fused into the check: one process per GPU, no communicator call in the step.
//
// Every rank owns one buffer of the same layout (32-bit words), mapped into every other rank's address
// space (torch.distributed._symmetric_memory, cudaIpc*, or one process with peer access):
//     [2][world][words]   two gather areas, used alternately (call parity), slot r = rank r's bitmask
//     [world]             signal row: word q = number of the last call whose slot rank q has delivered
//     [4]                 header: [0] number of the last call this rank has pushed (a fresh context
//                         picks the numbering up from here), [1] waits that timed out (sticky: the
//                         caller reads it with rc_gather_status), [2..3] reserved
// The two check kernels write the bitmask words straight into the rank's own slot.  The LAST CTA of
// the descent kernel to finish (the one that already resets the work-list counters) then runs
// gather_tail: it stores the slot into every peer's buffer with plain peer stores, fences, releases
// the call number into its word of every peer's signal row (st.release.sys) and -- unless the caller
// asked for the two-phase form -- acquires the peers' call numbers from its own row
// (ld.acquire.sys).  No third launch, no communicator.  Two areas, because a fast peer may already
// deliver call k+1 while this rank's stream still reads the result of call k; it cannot get to k+2
// before this rank has delivered k+1.  A peer that never arrives makes the wait give up after
// timeout_ns: counted in the header word and in rc_stats.watchdog, never a hang.
#pragma once

namespace rc {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Wait (thread p < world, p != rank) until rank p's call number `call` has arrived in the own row.
__device__ __forceinline__ void gather_wait_peers(const GatherArgs &a)
{
    const int p = (int)threadIdx.x;
    if (p < a.world && p != a.rank) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(a.my_row + p) - a.call) < 0) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > a.timeout_ns) {  // the peer never delivered: report, do not hang the device
                atomicAdd(a.my_row + a.world + 1, 1u);
                atomicAdd(a.counters + CNT_WATCHDOG, 1ULL);
                atomicOr(a.counters + CNT_WATCHDOG_WHERE, 8ULL);
                break;
            }
        }
    }
}

// All threads of ONE CTA; the slot `a.src` is complete and visible to this CTA.
__device__ __forceinline__ void gather_tail(const GatherArgs &a)
{
    for (int q = 1; q < a.world; q++) {
        uint32_t *dst = a.peer_slot[(a.rank + q) % a.world];
        for (int w = threadIdx.x; w < a.words; w += blockDim.x) dst[w] = __ldcg(a.src + w);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < a.world && (int)threadIdx.x != a.rank)
        st_release_sys(a.peer_signal[threadIdx.x], a.call);
    if (threadIdx.x == 0) a.my_row[a.world] = a.call;
    if (a.wait) gather_wait_peers(a);
    __syncthreads();
}

// Stand-alone forms: the push for calls whose shard is empty or not checked by the descent kernel,
// and the second phase of rc_check_gather_begin / rc_gather_end.
__global__ void __launch_bounds__(256) gather_push_kernel(const GatherArgs a)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the slot is written by the kernels before this one
    gather_tail(a);
}

__global__ void __launch_bounds__(32) gather_wait_kernel(const GatherArgs a)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    gather_wait_peers(a);
}

cudaError_t launch_gather_push(const rc_ctx *ctx, const GatherArgs &a, cudaStream_t stream)
{
    if (a.world <= 1) return cudaSuccess;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1u); cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ctx->knob_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, gather_push_kernel, a);
}

cudaError_t launch_gather_wait(const rc_ctx *ctx, const GatherArgs &a, cudaStream_t stream)
{
    if (a.world <= 1) return cudaSuccess;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1u); cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ctx->knob_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, gather_wait_kernel, a);
}

}  // namespace rc
