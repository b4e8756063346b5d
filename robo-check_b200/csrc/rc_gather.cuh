// rc_gather.cuh -- all-gather of the validity bitmask through peer-mapped memory (NVLink / NVSwitch),
// fused behind the check: one process per GPU, no communicator call in the step.
//
// Every rank owns one buffer of the same layout (32-bit words), mapped into every other rank's address
// space (torch.distributed._symmetric_memory, cudaIpc*, or one process with peer access):
//     [2][world][words]   two gather areas, used alternately (call parity), slot r = rank r's bitmask
//     [world]             signal row: word q = number of the last call whose slot rank q has delivered
// rc_check_gather runs the two check kernels with the bitmask going straight into the rank's own slot,
// then gather_push_kernel (programmatic dependent launch): CTA b serves peer p = (rank + 1 + b) % world,
// stores the slot into p's buffer with plain peer stores, fences, releases signal[rank] in p's row with
// the call number, and then waits until p's call number shows up in its own row.  When the kernel ends,
// the rank's gather area of this parity holds the bitmask of all ranks.  Two areas, because a fast peer
// may already deliver call k+1 while this rank's stream still reads the result of call k; it cannot get
// to k+2 before this rank has delivered k+1.  Every wait is bounded by the nanosecond timer: a peer that
// never arrives is counted in rc_stats.watchdog instead of hanging the device.
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

__global__ void __launch_bounds__(256) gather_push_kernel(const GatherArgs a)
{
    // the slot is written by the two kernels before this one in the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int p = (a.rank + 1 + (int)blockIdx.x) % a.world;
    uint32_t *dst = a.peer_slot[p];
    for (int w = threadIdx.x; w < a.words; w += blockDim.x) dst[w] = __ldcg(a.src + w);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        st_release_sys(a.peer_signal[p], a.call);
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(a.my_signal + p) - a.call) < 0) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > a.timeout_ns) {  // the peer never delivered: do not hang the device
                atomicAdd(a.counters + CNT_WATCHDOG, 1ULL);
                atomicOr(a.counters + CNT_WATCHDOG_WHERE, 8ULL);
                break;
            }
        }
    }
    __syncthreads();
}

cudaError_t launch_gather_push(const rc_ctx *ctx, const GatherArgs &a, cudaStream_t stream)
{
    if (a.world <= 1) return cudaSuccess;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.world - 1)); cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ctx->knob_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, gather_push_kernel, a);
}

}  // namespace rc
