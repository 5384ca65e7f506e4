// CenteredGrid (collocated) velocities: fluid.make_incompressible with wide_stencil = True (phi/physics/fluid.py:154-155, 197-202;
// SURVEY.md Appendix A "CenteredGrid velocity variant"; reference tests tests/commit/physics/test_fluid.py:34-36).
//
//   div     = sum_c (v_c[i+1] - v_c[i-1]) / (2 dx_c)            ghosts of v_c from the velocity boundary   (_field_math.py:627-632)
//   A p     = divergence_centered(gradient_centered(p))           gradient ghosts: pressure boundary; divergence ghosts: velocity
//                                                                 boundary with constants removed               (fluid.py:197-202)
//   v_c    -= (p[i+1] - p[i-1]) / (2 dx_c)                                                                  (fluid.py:158-161)
//
// The wide operator ([1 0 -2 0 1] / 4dx^2 in the interior) is not symmetric at the boundary rows, so the reference's default
// Solve() = 'auto' = CG-adaptive (PhiML/phiml/backend/_linalg.py:93-128) is what converges on it; that is the solver here, with
// the reference's rank-1 matrix_offset for rank-deficient systems (linear(), _linalg.py:784-789).
//
// This is NOT the tuned path (the north-star workloads are staggered): one thread per cell, four launches per iteration, dot
// products through double atomics, the host reads the per-entry status every few iterations (the entry point is a `_host` call).
// It exists so that CenteredGrid velocities run on the GPU with the reference's semantics instead of falling through.
#include "phi_internal.cuh"
#include "launch.cuh"

struct CoVec { DField f[3]; const float* p[3]; };          // three centred arrays with their own boundaries
struct CoOut { float* p[3]; };

struct CoStatus {            // per batch entry, device memory
    double dx_dy, dx_r, s_dx, rsq, r_dy;     // accumulators (zeroed by the kernel that consumes them last)
    float tol_sq, rsq0, last_rsq, pad_;
    int iterations, cont, converged, diverged;
};

template <int DIM>
__device__ __forceinline__ bool co_index(const DGrid& g, int& b, int& x, int& y, int& z)
{
    x = blockIdx.x * blockDim.x + threadIdx.x;
    y = blockIdx.y;
    const int zb = blockIdx.z;
    if (DIM == 3) { z = zb % g.n[2]; b = zb / g.n[2]; } else { z = 0; b = zb; }
    return x < g.n[0] && y < g.n[1];
}

__device__ __forceinline__ void co_block_add(double* dst, double v)
{
    // warp shuffle reduction, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v);
}

// g_c = (p[i + e_c] - p[i - e_c]) / (2 dx_c), ghosts from pf.   SUB: v_c -= g_c (final correction) instead of storing g_c.
template <int DIM, bool SUB>
__global__ void __launch_bounds__(128)
k_co_gradient(const __grid_constant__ DGrid g, const __grid_constant__ DField pf, const float* __restrict__ p, const __grid_constant__ CoOut out,
              const CoStatus* __restrict__ st)
{
    int b, x, y, z;
    const bool in = co_index<DIM>(g, b, x, y, z);
    if (!in || (st && !st[b].cont)) return;
    const long long off = (long long)b * pf.sb + (long long)z * pf.sz + (long long)y * pf.sy + x;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        const float up = phi_fetch<DIM>(p, g, pf, b, x + (c == 0), y + (c == 1), z + (c == 2));
        const float lw = phi_fetch<DIM>(p, g, pf, b, x - (c == 0), y - (c == 1), z - (c == 2));
        const float grad = phi_div(up - lw, g.dx[c] * 2.f, g.inv_dx[c] * 0.5f);
        if (SUB) out.p[c][off] = out.p[c][off] - grad; else out.p[c][off] = grad;
    }
}

// out = sum_c (v_c[i + e_c] - v_c[i - e_c]) / (2 dx_c), ghosts from v.f[c].
// MODE 0: plain (right-hand side).  MODE 1: out = A dir; accumulates dir.out, dir.r and sum(dir) (operator application in the loop).
template <int DIM, int MODE>
__global__ void __launch_bounds__(128)
k_co_divergence(const __grid_constant__ DGrid g, const __grid_constant__ CoVec v, const __grid_constant__ DField cf, float* __restrict__ out,
                const float* __restrict__ dir, const float* __restrict__ r, CoStatus* __restrict__ st)
{
    int b, x, y, z;
    const bool in = co_index<DIM>(g, b, x, y, z);
    double a0 = 0, a1 = 0, a2 = 0;
    const bool live = in && (MODE == 0 || st[b].cont);
    if (live) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            const float up = phi_fetch<DIM>(v.p[c], g, v.f[c], b, x + (c == 0), y + (c == 1), z + (c == 2));
            const float lw = phi_fetch<DIM>(v.p[c], g, v.f[c], b, x - (c == 0), y - (c == 1), z - (c == 2));
            const float term = phi_div(up - lw, g.dx[c] * 2.f, g.inv_dx[c] * 0.5f);
            acc = (c == 0) ? term : acc + term;
        }
        const long long off = (long long)b * cf.sb + (long long)z * cf.sz + (long long)y * cf.sy + x;
        out[off] = acc;
        if (MODE == 1) { const float d = dir[off]; a0 = (double)d * acc; a1 = (double)d * r[off]; a2 = d; }
    }
    if (MODE == 1) {
        // all threads of a block share b (one grid line per block row): reduce per warp
        const int bb = (DIM == 3) ? blockIdx.z / g.n[2] : blockIdx.z;
        co_block_add(&st[bb].dx_dy, a0); co_block_add(&st[bb].dx_r, a1); co_block_add(&st[bb].s_dx, a2);
    }
}

// Element-wise pieces of the CG-adaptive iteration (_linalg.py:109-122); q = A dir + c * sum(dir) with c = matrix_offset.
//   PHASE 0 (initial residual): r = y - q(x0); dir = r; accumulates |r|^2 and |y|^2
//   PHASE 1: step = (dir.r) / (dir.q); x += step dir; r -= step q; accumulates |r|^2 and r.q
//   PHASE 2: dir = r - ((r.q) / (dir.q)) dir; block 0 of every batch entry advances the iteration count and the stopping rule
template <int DIM, int PHASE>
__global__ void __launch_bounds__(128)
k_co_update(const __grid_constant__ DGrid g, const __grid_constant__ DField cf, float* __restrict__ x, float* __restrict__ r, float* __restrict__ dir,
            const float* __restrict__ q, const float* __restrict__ y, float offset, CoStatus* __restrict__ st,
            const float* __restrict__ means, const double* __restrict__ x0sums)
{
    int b, xx, yy, zz;
    const bool in = co_index<DIM>(g, b, xx, yy, zz);
    const int bb = (DIM == 3) ? blockIdx.z / g.n[2] : blockIdx.z;
    const long long off = (long long)bb * cf.sb + (long long)zz * cf.sz + (long long)yy * cf.sy + xx;
    CoStatus& s = st[bb];
    double a0 = 0, a1 = 0;
    if (PHASE == 0) {
        if (in) {
            const float yv = y[off] - (means ? means[bb] : 0.f);
            const float rv = yv - (q[off] + offset * (float)x0sums[bb]);
            r[off] = rv; dir[off] = rv;
            a0 = (double)rv * rv; a1 = (double)yv * yv;
        }
        co_block_add(&s.rsq, a0); co_block_add(&s.r_dy, a1);          // r_dy doubles as |y|^2 during set-up
        return;
    }
    if (!s.cont) return;
    const double S = s.s_dx;
    const double dxdy = s.dx_dy + (double)offset * S * S;               // dir . (A dir + c sum(dir))
    if (PHASE == 1) {
        const float step = dxdy != 0.0 ? (float)(s.dx_r / dxdy) : 0.f;  // divide_no_nan
        if (in) {
            const float qv = q[off] + offset * (float)S;
            x[off] = x[off] + step * dir[off];
            const float rv = r[off] - step * qv;
            r[off] = rv;
            a0 = (double)rv * rv; a1 = (double)rv * qv;
        }
        co_block_add(&s.rsq, a0); co_block_add(&s.r_dy, a1);
        return;
    }
    // PHASE 2
    const float coef = dxdy != 0.0 ? (float)(s.r_dy / dxdy) : 0.f;
    if (in) dir[off] = r[off] - coef * dir[off];
}

// one thread per batch entry, between the phases: bookkeeping of the accumulators and of the stopping rule (stop_on_l2)
__global__ void k_co_control(CoStatus* st, int batch, int phase, float rtol, float atol, int max_iter)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    CoStatus& s = st[b];
    if (phase == 0) {                 // after the initial residual: tolerance relative to |y|^2 (_linalg.py:109)
        const float rsq = (float)s.rsq;
        s.tol_sq = fmaxf(rtol * rtol * (float)s.r_dy, atol * atol);
        s.rsq0 = fabsf(rsq);
        s.converged = fabsf(rsq) <= s.tol_sq; s.diverged = !isfinite(rsq);
        s.iterations = 0;
        s.cont = (!s.converged && !s.diverged && max_iter > 0) ? 1 : 0;
        s.rsq = 0; s.r_dy = 0; s.dx_dy = 0; s.dx_r = 0; s.s_dx = 0;
        return;
    }
    if (!s.cont) return;
    if (phase == 1) { s.rsq = 0; s.r_dy = 0; return; }              // before PHASE 1 accumulates
    // phase 2: after the direction update - the operator application that follows refills dx_dy, dx_r, s_dx
    const float rsq = fabsf((float)s.rsq);
    s.iterations += 1;
    s.converged = rsq <= s.tol_sq;
    s.diverged = !isfinite(rsq) || (rsq / s.rsq0 > 1e5f && s.iterations >= 8);
    s.cont = (!s.converged && !s.diverged && s.iterations < max_iter) ? 1 : 0;
    s.last_rsq = rsq;                  // residual_sq for the result record
    s.dx_dy = 0; s.dx_r = 0; s.s_dx = 0;
}

__global__ void k_co_mean(const __grid_constant__ DGrid g, const __grid_constant__ DField cf, const float* __restrict__ a, double* __restrict__ sums)
{
    int b, x, y, z;
    const bool in = g.dim == 3 ? co_index<3>(g, b, x, y, z) : co_index<2>(g, b, x, y, z);
    const int bb = (g.dim == 3) ? blockIdx.z / g.n[2] : blockIdx.z;
    double v = 0;
    if (in) v = a[(long long)bb * cf.sb + (long long)z * cf.sz + (long long)y * cf.sy + x];
    co_block_add(&sums[bb], v);
}

__global__ void k_co_finish(const CoStatus* st, PhiCgResult* result, const double* sums, float* means, double cells, int batch, int what)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (what == 0) { means[b] = (float)(sums[b] / cells); return; }
    PhiCgResult res;
    res.iterations = st[b].iterations; res.converged = st[b].converged; res.diverged = st[b].diverged;
    res.residual_sq = st[b].iterations > 0 ? st[b].last_rsq : st[b].rsq0; res.tol_sq = st[b].tol_sq; res.initial_residual_sq = st[b].rsq0;
    result[b] = res;
}

template <int DIM>
__global__ void __launch_bounds__(128)
k_co_sub_mean(const __grid_constant__ DGrid g, const __grid_constant__ DField cf, float* __restrict__ a, const float* __restrict__ means)
{
    int b, x, y, z;
    if (!co_index<DIM>(g, b, x, y, z)) return;
    a[(long long)b * cf.sb + (long long)z * cf.sz + (long long)y * cf.sy + x] -= means[b];
}

static dim3 co_grid(const DGrid& g) { return dim3((g.n[0] + 127) / 128, g.n[1], g.n[2] * g.batch); }

size_t phi_collocated_workspace_bytes(const DGrid& g)
{
    const size_t arr = ((size_t)g.cext[0] * g.cext[1] * g.cext[2] * g.batch * sizeof(float) + 255) / 256 * 256;
    // r, dir, q, div, 3 gradient components + status + sums of div / of x0 + means
    return 7 * arr + (size_t)g.batch * (sizeof(CoStatus) + 2 * sizeof(double) + sizeof(float)) + 1024;
}

// Host-synchronising: reads the per-entry status every `poll` iterations.
int phi_make_incompressible_collocated(const DGrid& g, const DField vfields[3], const DField vfields0[3], const DField& pf, const DField& cf,
                                       float* const v[3], float* p, const PhiCgParams& prm, int balance, PhiCgResult* result,
                                       void* workspace, size_t ws_bytes, cudaStream_t s)
{
    if (ws_bytes < phi_collocated_workspace_bytes(g)) { phi_set_error("collocated: workspace %zu < %zu bytes", ws_bytes, phi_collocated_workspace_bytes(g)); return PHI_ERR_WORKSPACE; }
    const size_t arr = ((size_t)g.cext[0] * g.cext[1] * g.cext[2] * g.batch * sizeof(float) + 255) / 256 * 256;
    unsigned char* ws = (unsigned char*)workspace;
    float* r = (float*)ws; float* dir = (float*)(ws + arr); float* q = (float*)(ws + 2 * arr); float* div = (float*)(ws + 3 * arr);
    CoOut grad; for (int c = 0; c < 3; ++c) grad.p[c] = (float*)(ws + (4 + c) * arr);
    CoStatus* st = (CoStatus*)(ws + 7 * arr);
    double* sums = (double*)(st + g.batch);
    double* xsums = sums + g.batch;
    float* means = (float*)(xsums + g.batch);
    const int B = g.batch, tb = (B + 63) / 64;
    const dim3 grid = co_grid(g), block(128);
    const double cells = (double)g.n[0] * g.n[1] * g.n[2];
    cudaError_t e = cudaMemsetAsync(st, 0, (size_t)B * (sizeof(CoStatus) + 2 * sizeof(double) + sizeof(float)), s);
    if (e) return (int)e;
    CoVec vin, gvec;
    for (int c = 0; c < 3; ++c) { vin.f[c] = vfields[c]; vin.p[c] = v[c]; gvec.f[c] = vfields0[c]; gvec.p[c] = grad.p[c]; }
    const bool d3 = g.dim == 3;
#define CO_LAUNCH(K2, K3, ...) do { if (d3) K3<<<grid, block, 0, s>>>(__VA_ARGS__); else K2<<<grid, block, 0, s>>>(__VA_ARGS__); } while (0)
    // right-hand side: divergence of the input velocity, balanced when the system is rank deficient (fluid.py:145-148, 205-209)
    CO_LAUNCH((k_co_divergence<2, 0>), (k_co_divergence<3, 0>), g, vin, cf, div, nullptr, nullptr, st);
    if (balance) {
        k_co_mean<<<grid, block, 0, s>>>(g, cf, div, sums);
        k_co_finish<<<tb, 64, 0, s>>>(st, result, sums, means, cells, B, 0);
    }
    auto apply = [&](const float* vec, bool loop) {          // q = A vec (and, inside the loop, the dot products)
        CO_LAUNCH((k_co_gradient<2, false>), (k_co_gradient<3, false>), g, pf, vec, grad, loop ? st : nullptr);
        if (loop) CO_LAUNCH((k_co_divergence<2, 1>), (k_co_divergence<3, 1>), g, gvec, cf, q, vec, r, st);
        else      CO_LAUNCH((k_co_divergence<2, 0>), (k_co_divergence<3, 0>), g, gvec, cf, q, nullptr, nullptr, st);
    };
    // r0 = y - (A + c 11^T) x0
    if (prm.matrix_offset != 0.f) k_co_mean<<<grid, block, 0, s>>>(g, cf, p, xsums);
    apply(p, false);
    CO_LAUNCH((k_co_update<2, 0>), (k_co_update<3, 0>), g, cf, p, r, dir, q, div, prm.matrix_offset, st, balance ? means : nullptr, xsums);
    k_co_control<<<tb, 64, 0, s>>>(st, B, 0, prm.rtol, prm.atol, prm.max_iter);
    apply(dir, true);
    int h_cont = 1;
    CoStatus* hst = (CoStatus*)malloc(sizeof(CoStatus) * B);
    if (!hst) { phi_set_error("collocated: out of host memory"); return PHI_ERR_INVALID; }
    for (int it = 0; it < prm.max_iter && h_cont; ++it) {
        k_co_control<<<tb, 64, 0, s>>>(st, B, 1, prm.rtol, prm.atol, prm.max_iter);
        CO_LAUNCH((k_co_update<2, 1>), (k_co_update<3, 1>), g, cf, p, r, dir, q, div, prm.matrix_offset, st, nullptr, xsums);
        CO_LAUNCH((k_co_update<2, 2>), (k_co_update<3, 2>), g, cf, p, r, dir, q, div, prm.matrix_offset, st, nullptr, xsums);
        k_co_control<<<tb, 64, 0, s>>>(st, B, 2, prm.rtol, prm.atol, prm.max_iter);
        apply(dir, true);
        if ((it & 7) == 7 || it + 1 == prm.max_iter) {
            e = cudaMemcpyAsync(hst, st, sizeof(CoStatus) * B, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e) { free(hst); return (int)e; }
            h_cont = 0;
            for (int b = 0; b < B; ++b) h_cont |= hst[b].cont;
        }
    }
    free(hst);
    k_co_finish<<<tb, 64, 0, s>>>(st, result, sums, means, cells, B, 1);
    // v -= grad p
    CoOut vout; for (int c = 0; c < 3; ++c) vout.p[c] = v[c];
    CO_LAUNCH((k_co_gradient<2, true>), (k_co_gradient<3, true>), g, pf, p, vout, nullptr);
#undef CO_LAUNCH
    return (int)cudaGetLastError();
}

// y = divergence_centered(gradient_centered(x)): fluid.masked_laplace(wide_stencil=True) on its own (matrix_offset estimate, tests)
int phi_wide_laplace(const DGrid& g, const DField vfields0[3], const DField& pf, const DField& cf, const float* x, float* y,
                     void* workspace, size_t ws_bytes, cudaStream_t s)
{
    const size_t arr = ((size_t)g.cext[0] * g.cext[1] * g.cext[2] * g.batch * sizeof(float) + 255) / 256 * 256;
    if (ws_bytes < 3 * arr) { phi_set_error("wide_laplace: workspace %zu < %zu bytes", ws_bytes, 3 * arr); return PHI_ERR_WORKSPACE; }
    CoOut grad; CoVec gvec;
    for (int c = 0; c < 3; ++c) { grad.p[c] = (float*)((unsigned char*)workspace + c * arr); gvec.f[c] = vfields0[c]; gvec.p[c] = grad.p[c]; }
    const dim3 grid = co_grid(g), block(128);
    if (g.dim == 3) {
        k_co_gradient<3, false><<<grid, block, 0, s>>>(g, pf, x, grad, nullptr);
        k_co_divergence<3, 0><<<grid, block, 0, s>>>(g, gvec, cf, y, nullptr, nullptr, nullptr);
    } else {
        k_co_gradient<2, false><<<grid, block, 0, s>>>(g, pf, x, grad, nullptr);
        k_co_divergence<2, 0><<<grid, block, 0, s>>>(g, gvec, cf, y, nullptr, nullptr, nullptr);
    }
    return (int)cudaGetLastError();
}
