// mb_peaks.cu -- measured per-GPU peaks for the rooflines of kernels that are NOT HBM-bound:
// fp64-pipe rate, integer-ALU / FMA issue rates, L2 read bandwidth, L1 hit bandwidth, shared-memory bandwidth.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/mb_peaks profiles/tools/mb_peaks.cu
//   profiles/tools/mb_peaks > profiles/peaks.json        (on the B200)
//
// Every kernel is a full-chip launch (148 SMs x resident CTAs), timed with CUDA events, best of 5 after a warm-up.
// Rates are per GPU: lane-instructions per second for the pipes, bytes per second for the memories.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int OP>
__global__ void __launch_bounds__(256) k_fp64(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { // DFMA
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        } else if (OP == 1) { // DADD
            x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b);
            x4 = __dadd_rn(x4, b); x5 = __dadd_rn(x5, b); x6 = __dadd_rn(x6, b); x7 = __dadd_rn(x7, b);
        } else { // DMUL
            x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a);
            x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int OP>
__global__ void __launch_bounds__(256) k_int(unsigned *out, int iters, unsigned a, unsigned b) {
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float f0 = x0, f1 = x1, f2 = x2, f3 = x3, f4 = x4, f5 = x5, f6 = x6, f7 = x7;
    const float fa = __uint_as_float(a), fb = __uint_as_float(b);
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { // IMAD
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b; x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        } else if (OP == 1) { // LOP3 (integer ALU)
            x0 = (x0 ^ a) & (x0 | b); x1 = (x1 ^ a) & (x1 | b); x2 = (x2 ^ a) & (x2 | b); x3 = (x3 ^ a) & (x3 | b);
            x4 = (x4 ^ a) & (x4 | b); x5 = (x5 ^ a) & (x5 | b); x6 = (x6 ^ a) & (x6 | b); x7 = (x7 ^ a) & (x7 | b);
        } else { // FFMA
            f0 = fmaf(f0, fa, fb); f1 = fmaf(f1, fa, fb); f2 = fmaf(f2, fa, fb); f3 = fmaf(f3, fa, fb);
            f4 = fmaf(f4, fa, fb); f5 = fmaf(f5, fa, fb); f6 = fmaf(f6, fa, fb); f7 = fmaf(f7, fa, fb);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
}

// L2 read bandwidth: every CTA streams the whole buffer (resident in L2) with 16-byte ld.global.cg loads (no L1)
__global__ void __launch_bounds__(256) k_l2(const uint4 *buf, size_t n16, int passes, unsigned *out) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) out[0] = acc;
}

// L1 hit bandwidth: a CTA re-reads its own 32 KiB window with 16-byte read-only loads
__global__ void __launch_bounds__(256) k_l1(const uint4 *buf, int iters, unsigned *out) {
    const uint4 *w = buf + (size_t)blockIdx.x * 2048; // 32 KiB per CTA
    unsigned acc = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 v = __ldg(w + ((threadIdx.x + 256 * k + 37 * i) & 2047));
            acc += v.x ^ v.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// shared-memory bandwidth: conflict-free 16-byte loads (LDS.128), 32 KiB per CTA
__global__ void __launch_bounds__(256) k_smem(int iters, unsigned *out) {
    __shared__ uint4 s[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned acc = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 v = s[(threadIdx.x + 256 * k + 8 * i) & 2047];
            acc += v.x ^ v.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// shared-memory broadcast reads (all lanes of a warp read the same 16 bytes): what the fused kernel's scan does
__global__ void __launch_bounds__(256) k_smem_bcast(int iters, unsigned *out) {
    __shared__ uint4 s[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned acc = 0;
    const int w = threadIdx.x >> 5;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 v = s[(w * 64 + k + 8 * i) & 2047];
            acc += v.x ^ v.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static double best_ms(F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const int ctas = sms * 8, threads = 256; // 2048 threads per SM
    double *dout; unsigned *uout;
    CK(cudaMalloc(&dout, (size_t)ctas * threads * sizeof(double)));
    CK(cudaMalloc(&uout, (size_t)ctas * threads * sizeof(unsigned)));
    const int it = 1 << 14;
    const double lanes = (double)ctas * threads * 8.0 * it;
    const double t_dfma = best_ms([&] { k_fp64<0><<<ctas, threads>>>(dout, it, 1.0000001, 1e-9); });
    const double t_dadd = best_ms([&] { k_fp64<1><<<ctas, threads>>>(dout, it, 1.0000001, 1e-9); });
    const double t_dmul = best_ms([&] { k_fp64<2><<<ctas, threads>>>(dout, it, 1.0000001, 1e-9); });
    const int iti = 1 << 16;
    const double lanesi = (double)ctas * threads * 8.0 * iti;
    const double t_imad = best_ms([&] { k_int<0><<<ctas, threads>>>(uout, iti, 1664525u, 1013904223u); });
    const double t_lop = best_ms([&] { k_int<1><<<ctas, threads>>>(uout, iti, 0x9E3779B9u, 0x7F4A7C15u); }); // one LOP3 per statement (3-input LUT)
    const double t_ffma = best_ms([&] { k_int<2><<<ctas, threads>>>(uout, iti, 0x3F800001u, 0x33000000u); });
    // L2: 32 MiB buffer, warmed into L2 by the warm-up launch
    const size_t l2_bytes = 32u << 20;
    uint4 *buf;
    CK(cudaMalloc(&buf, l2_bytes));
    CK(cudaMemset(buf, 1, l2_bytes));
    const int passes = 64;
    const double t_l2 = best_ms([&] { k_l2<<<sms * 8, 256>>>(buf, l2_bytes / 16, passes, uout); });
    const int itm = 1 << 12;
    const double t_l1 = best_ms([&] { k_l1<<<sms * 4, 256>>>(buf, itm, uout); });
    const double t_sm = best_ms([&] { k_smem<<<sms * 4, 256>>>(itm, uout); });
    const double t_sb = best_ms([&] { k_smem_bcast<<<sms * 4, 256>>>(itm, uout); });
    const double mem_bytes = (double)sms * 4 * 256 * 8.0 * itm * 16.0;
    printf("{\n");
    printf(" \"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_nominal\": %.0f,\n", prop.name, sms, clock_khz / 1e3);
    printf(" \"fp64_dfma_lane_inst_per_s\": %.4e, \"fp64_dadd_lane_inst_per_s\": %.4e, \"fp64_dmul_lane_inst_per_s\": %.4e,\n",
           lanes / (t_dfma * 1e-3), lanes / (t_dadd * 1e-3), lanes / (t_dmul * 1e-3));
    printf(" \"imad_lane_inst_per_s\": %.4e, \"lop3_lane_inst_per_s\": %.4e, \"ffma_lane_inst_per_s\": %.4e,\n",
           lanesi / (t_imad * 1e-3), lanesi / (t_lop * 1e-3), lanesi / (t_ffma * 1e-3));
    printf(" \"issue_warp_inst_per_s_nominal\": %.4e,\n", (double)sms * 4 * clock_khz * 1e3);
    printf(" \"l2_read_GBps\": %.1f, \"l1_hit_read_GBps\": %.1f, \"smem_read_GBps\": %.1f, \"smem_broadcast_warp_loads_per_s\": %.4e,\n",
           (double)l2_bytes * passes / (t_l2 * 1e-3) / 1e9, mem_bytes / (t_l1 * 1e-3) / 1e9,
           mem_bytes / (t_sm * 1e-3) / 1e9, (double)sms * 4 * 8 * 8.0 * itm / (t_sb * 1e-3));
    printf(" \"how\": \"profiles/tools/mb_peaks.cu: full-chip launches, 8 independent chains per thread, CUDA events, best of 5\"\n");
    printf("}\n");
    return 0;
}
