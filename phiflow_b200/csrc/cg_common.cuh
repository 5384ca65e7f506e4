// Shared pieces of the two CG implementations (LDG marching kernel in cg_kernels.cu, TMA ring kernel in ring_kernels.cu).
#pragma once
#include "phi_internal.cuh"

#define CG_MAX_BATCH 1024
#define CG_MAX_GRID 2048

struct CgArgs {
    DGrid g; DField pf; UnitMap um;
    const float* rhs; float* x; float* r; float* d0; float* d1;
    const float* acc;              // N4: accessible mask (1 = fluid, 0 = obstacle) or nullptr
    double* partials;              // [2 regions][2 accumulators][batch][grid]
    PhiCgResult* result;
    PhiCgParams prm;
};

struct CgShared {                   // per-thread view of the CTA's shared memory (pointers carved from one dynamic block)
    double (*warp_acc)[32];   // [2][warps] (up to 32 warps per CTA)
    double* sum0;                  // reduced accumulator 0 per batch entry
    double* sum1;
    double* delta;
    float* alpha;
    float* aprev;                  // alpha of the previous iteration (deferred x update of the ring kernel)
    float* beta;
    float* offs;                   // c * S  (offset part of q)
    float* mean;
    float* tol_sq;
    float* rsq0;
    int* iters;
    int* any_cont;
    unsigned char *cont, *conv, *divg;
};

__host__ __device__ inline size_t cg_smem_bytes(int batch)
{
    const size_t b8 = ((size_t)batch + 1) / 2 * 2;      // keep 8-byte alignment of what follows
    return 2 * 32 * sizeof(double) + 3 * b8 * sizeof(double) + 7 * b8 * sizeof(float)
         + (b8 + 2) * sizeof(int) + 3 * (b8 + 16);
}

__device__ __forceinline__ CgShared cg_carve(unsigned char* base, int batch)
{
    const size_t b8 = ((size_t)batch + 1) / 2 * 2;
    CgShared sh;
    unsigned char* p = base;
    sh.warp_acc = reinterpret_cast<double (*)[32]>(p); p += 2 * 32 * sizeof(double);
    sh.sum0 = (double*)p; p += b8 * sizeof(double);
    sh.sum1 = (double*)p; p += b8 * sizeof(double);
    sh.delta = (double*)p; p += b8 * sizeof(double);
    sh.alpha = (float*)p; p += b8 * sizeof(float);
    sh.aprev = (float*)p; p += b8 * sizeof(float);
    sh.beta = (float*)p; p += b8 * sizeof(float);
    sh.offs = (float*)p; p += b8 * sizeof(float);
    sh.mean = (float*)p; p += b8 * sizeof(float);
    sh.tol_sq = (float*)p; p += b8 * sizeof(float);
    sh.rsq0 = (float*)p; p += b8 * sizeof(float);
    sh.iters = (int*)p; p += b8 * sizeof(int);
    sh.any_cont = (int*)p; p += 2 * sizeof(int);
    sh.cont = p; p += b8 + 16;
    sh.conv = p; p += b8;
    sh.divg = p;
    return sh;
}


#ifndef CG_BLOCK_WARPS
#define CG_BLOCK_WARPS PHI_WARPS_PER_CTA
#endif

// Block-level flush of the two fp32 thread accumulators of the current batch entry into partials[region][k][b][cta].
__device__ __forceinline__ void flush_partials(const CgShared& sh, double* partials, int region, int batch, int b, double a0, double a1)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if (lane == 0) { sh.warp_acc[0][warp] = a0; sh.warp_acc[1][warp] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s0 += sh.warp_acc[0][w]; s1 += sh.warp_acc[1][w]; }
        const size_t G = gridDim.x;
        partials[((size_t)(region * 2 + 0) * batch + b) * G + blockIdx.x] = s0;
        partials[((size_t)(region * 2 + 1) * batch + b) * G + blockIdx.x] = s1;
    }
    __syncthreads();
}

// After a grid barrier: every CTA sums, in a fixed order, the partials of the CTAs that own units of batch entry b.
__device__ __forceinline__ void reduce_partials(const CgShared& sh, const double* partials, int region, int batch, int units_per_batch,
                                                const unsigned char* active)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    const int cnt = min(units_per_batch, G);
    for (int b = warp; b < batch; b += (int)(blockDim.x >> 5)) {
        if (active && !active[b]) continue;
        const int first = (int)(((long long)b * units_per_batch) % G);
        double s0 = 0, s1 = 0;
        for (int i = lane; i < cnt; i += 32) {
            int c = first + i; if (c >= G) c -= G;
            s0 += __ldcg(&partials[((size_t)(region * 2 + 0) * batch + b) * G + c]);
            s1 += __ldcg(&partials[((size_t)(region * 2 + 1) * batch + b) * G + c]);
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sh.sum0[b] = s0; sh.sum1[b] = s1; }
    }
    __syncthreads();
}

