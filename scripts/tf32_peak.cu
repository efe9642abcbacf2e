// tf32_peak.cu -- the measured tcgen05.mma kind::tf32 rate on this pool's B200 (BASELINE.md section 2 / SURVEY section 6 ask
// for it as a kept artefact; MEASURED_PEAKS.json only has the cuBLAS bf16 figure).  A bare UMMA loop: every CTA (one per
// SM, 148) issues back-to-back 128 x 256 x 8 tf32 instructions from one thread into two alternating TMEM accumulators,
// operands resident on chip, nothing else running:
//   form 0  SS: A (128 x 32 floats) and B (256 x 32 floats) read from shared memory (12 KB per instruction)
//   form 1  TS: A copied once to tensor memory (tcgen05.cp), only B read from shared memory (8 KB per instruction)
// Prints one JSON line per form.  Build + run: scripts/tf32_peak.py (nvcc -arch sm_100a, the same wrappers as the product).
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace erp;

namespace erp { void set_error(const char*, ...) {} }     // tc_common.cuh's host helpers report through it; unused here

constexpr int A_BYTES = 2 * 128 * 128, B_BYTES = 2 * 256 * 128;      // two 32-float k chunks of each operand (K = 64, the product's tile)
static uint32_t idesc_for(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__global__ void __launch_bounds__(128, 1) umma_loop(int iters, int form, int n, int ksteps, uint32_t IDESC, float* sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + A_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_smem + B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    // pseudo-random tf32 operands (zeros would understate the power the pipe draws)
    uint32_t s = 1234567u + blockIdx.x * 977u + threadIdx.x;
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        s = s * 1664525u + 1013904223u;
        reinterpret_cast<float*>(smem)[i] = tf32_rna((float)((s >> 8) & 0xffff) / 65536.0f - 0.5f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(a_smem), b = smem_u32(b_smem);
        // form 1: the A tile lives in TMEM columns [480, 512): four k steps of 8 columns (re-used for every k step)
        const uint32_t a_tmem = tmem + 480;
        if (form == 1)
            for (int k = 0; k < 4; k++) tc_cp_128x256b(a_tmem + k * 8, smem_desc_sw128(a + k * 32));
        // `ksteps` instructions accumulate into one accumulator before the issuer moves to the other one, as a GEMM k loop does
        for (int it = 0; it < iters; it++) {
            const uint32_t d = tmem + (it & 1) * n;          // two disjoint accumulators of n columns
            for (int k = 0; k < ksteps; k++) {
                const uint32_t ao = (uint32_t)((k >> 2) & 1) * (A_BYTES / 2) + (k & 3) * 32, bo = (uint32_t)((k >> 2) & 1) * (B_BYTES / 2) + (k & 3) * 32;
                if (form == 0) tc_mma_tf32(d, smem_desc_sw128(a + ao), smem_desc_sw128(b + bo), IDESC, k != 0 || it > 1);
                else tc_mma_tf32_ts(d, a_tmem + (k & 3) * 8, smem_desc_sw128(b + bo), IDESC, k != 0 || it > 1);
            }
        }
        tc_commit(bar);
        mbar_wait(bar, 0);
        tc_fence_after();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v[8];
        tc_ld8(tmem + ((threadIdx.x & ~31u) << 16), v);
        tc_wait_ld8(v);
        if (sink) sink[blockIdx.x * 32 + threadIdx.x] = __uint_as_float(v[0]);
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main(int argc, char** argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 40000;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int smem = A_BYTES + B_BYTES + 64;
    cudaFuncSetAttribute(umma_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* sink;
    cudaMalloc(&sink, sizeof(float) * 32 * sms);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // N = 256 is the product's tile; with A in tensor memory two accumulators + A must fit 512 columns: N = 240.
    // SS is measured at both widths so that SS and TS compare like for like; k steps per accumulator visit: 4, 8 (the
    // product: 8 + the norm step), 16.
    const int forms[3] = {0, 0, 1}, widths[3] = {256, 240, 240}, steps[3] = {4, 8, 16};
    for (int c = 0; c < 3; c++)
        for (int ks = 0; ks < 3; ks++) {
            const int form = forms[c], n = widths[c], ksteps = steps[ks], its = iters * 4 / ksteps;
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                umma_loop<<<sms, 128, smem>>>(its, form, n, ksteps, idesc_for(n), sink);
                cudaEventRecord(e1);
                if (cudaEventSynchronize(e1) != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const double instr = (double)its * ksteps, flop = instr * 2.0 * 128 * n * 8 * sms;
            printf("{\"form\": \"%s\", \"n\": %d, \"ksteps\": %d, \"ms\": %.4f, \"tflops\": %.1f, \"clk_per_instr_at_1965MHz\": %.1f, \"sms\": %d, \"instr_per_cta\": %.0f}\n",
                   form == 0 ? "SS" : "TS", n, ksteps, best, flop / (best * 1e-3) / 1e12, best * 1e-3 * 1.965e9 / instr, sms, instr);
        }
    return 0;
}
