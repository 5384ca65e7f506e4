// Vectorised projection stencils (A5 divergence, A6 gradient subtraction) and the one-launch semi-Lagrangian advection
// (A9-A11) of the incompressible step, replacing the one-thread-per-sample kernels of round 1 (stencil_kernels.cu /
// advect_kernels.cu keep those as the obstacle-mask variants and as the arithmetic these kernels must reproduce).
//
// Design (all HBM-bound, SURVEY.md section 8d byte table):
//   * a warp owns one grid line (y, z) and walks a 128-cell x segment; everything that depends only on the line - which
//     neighbouring lines a stencil reads and whether the boundary turns them into wrapped lines, clamped lines or
//     constants - is resolved ONCE per warp (warp-uniform), not per access.  x neighbours come from warp shuffles, only the
//     two edge lanes of a segment load them (boundary resolved there).
//   * divergence / grad_sub: one float4 (4 cells) per thread, every operand a 16-byte load; the y/z neighbour lines are L1/L2
//     hits (eight consecutive lines per CTA, z planes one sweep apart: 1 MiB per array and plane against 126 MB of L2), so
//     DRAM traffic is the compulsory 16 / 28 B per cell.
//   * advection: lanes own consecutive x (coalesced gathers for smooth displacement fields).  The three staggered components
//     are advected in ONE launch: the 11 velocity lines a cell's three faces need (shift resampling,
//     phi/field/_resample.py:341-364) are loaded once and shared, x-shifted values are shuffles; buoyancy
//     (resample(s * b, to=v), _resample.py:272-276) and the smoke inflow are epilogues of the same kernels, so the step
//     never re-reads a freshly written array just to add to it.
// Arithmetic (operation order, 0.5/0.5 lerp order of sample_subgrid, weighted 2^d sum of _ops.py:1010-1014) is identical to
// the scalar kernels, which are pinned against the oracle.
#include "phi_internal.cuh"
#include "launch.cuh"

#define FK_WARPS 8
#define FK_THREADS (FK_WARPS * 32)

// value at index x of a resolved line; x outside the stored range follows the boundary (same resolution order as phi_fetch)
__device__ __forceinline__ float fk_ldx(const float* __restrict__ a, const RowRef<3>& r, const DField& f, int x)
{
    if (x < f.lo[0] || x > f.hi[0]) { float c; if (!phi_resolve(x, f, 0, c)) return c; }
    if (r.off < 0) return r.cval;
    return __ldg(a + r.off + x);
}

template <int DIM>
__device__ __forceinline__ RowRef<3> fk_row(const DGrid& g, const DField& f, int b, int y, int z)
{
    RowRef<3> r; r.cval = 0.f; r.off = -1;
    if (!phi_resolve(y, f, 1, r.cval)) return r;
    if (DIM == 3) { if (!phi_resolve(z, f, 2, r.cval)) return r; } else z = 0;
    r.off = (long long)b * f.sb + (long long)z * f.sz + (long long)y * f.sy;
    return r;
}

// four consecutive values x0 .. x0+3 of a resolved line (x0 % 4 == 0).  Fast when all four are stored values.
__device__ __forceinline__ float4 fk_ld4(const float* __restrict__ a, const RowRef<3>& r, const DField& f, int x0)
{
    if (x0 >= f.lo[0] && x0 + 3 <= f.hi[0] && r.off >= 0) return __ldg(reinterpret_cast<const float4*>(a + r.off + x0));
    return make_float4(fk_ldx(a, r, f, x0), fk_ldx(a, r, f, x0 + 1), fk_ldx(a, r, f, x0 + 2), fk_ldx(a, r, f, x0 + 3));
}

struct FkLine { int b, y, z, x0; bool ok; };

// warp -> (line, 128-cell segment); lane -> float4 group.  Lines are numbered over the ALLOCATED extent so that upper
// boundary faces are covered.
template <int DIM>
__device__ __forceinline__ FkLine fk_line4(const DGrid& g)
{
    FkLine L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    L.x0 = (blockIdx.x * 32 + lane) * 4;
    L.y = blockIdx.y * FK_WARPS + warp;
    const int zb = blockIdx.z;
    if (DIM == 3) { L.z = zb % g.fext[2] - g.halo; L.b = zb / g.fext[2]; } else { L.z = 0; L.b = zb; }
    L.ok = L.y < g.fext[1];
    return L;
}

// ---------------------------------------------------------------------------------------------------------
// A5  divergence:  div = sum_d (v_d[i + e_d] - v_d[i]) / dx_d      16 B/cell (3-D), 12 (2-D)
// ---------------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(FK_THREADS)
k_div_vec(DGrid g, DVec v, DField cf, float* __restrict__ div)
{
    const FkLine L = fk_line4<DIM>(g);
    if (!L.ok || L.y >= g.n[1] || L.z < 0 || L.z >= g.n[2]) return;          // whole warp leaves together
    const int lane = threadIdx.x & 31;
    const int x0 = L.x0;
    const bool in_line = x0 < g.n[0];
    const RowRef<3> rx = fk_row<DIM>(g, v.f[0], L.b, L.y, L.z);
    const RowRef<3> ry0 = fk_row<DIM>(g, v.f[1], L.b, L.y, L.z), ry1 = fk_row<DIM>(g, v.f[1], L.b, L.y + 1, L.z);
    float4 ax = f4_splat(0.f), ay0 = ax, ay1 = ax, az0 = ax, az1 = ax;
    if (in_line) {
        ax = fk_ld4(v.p[0], rx, v.f[0], x0);
        ay0 = fk_ld4(v.p[1], ry0, v.f[1], x0); ay1 = fk_ld4(v.p[1], ry1, v.f[1], x0);
        if (DIM == 3) {
            const RowRef<3> rz0 = fk_row<DIM>(g, v.f[2], L.b, L.y, L.z), rz1 = fk_row<DIM>(g, v.f[2], L.b, L.y, L.z + 1);
            az0 = fk_ld4(v.p[2], rz0, v.f[2], x0); az1 = fk_ld4(v.p[2], rz1, v.f[2], x0);
        }
    }
    float nx = __shfl_down_sync(0xffffffffu, ax.x, 1);                      // v_x[x0 + 4]
    if (in_line && (lane == 31 || x0 + 4 >= g.n[0])) nx = fk_ldx(v.p[0], rx, v.f[0], x0 + 4);
    if (!in_line) return;
    const float dx = g.dx[0], dy = g.dx[1], dz = g.dx[2];
    float4 o;
    o.x = __fdiv_rn(ax.y - ax.x, dx) + __fdiv_rn(ay1.x - ay0.x, dy);
    o.y = __fdiv_rn(ax.z - ax.y, dx) + __fdiv_rn(ay1.y - ay0.y, dy);
    o.z = __fdiv_rn(ax.w - ax.z, dx) + __fdiv_rn(ay1.z - ay0.z, dy);
    o.w = __fdiv_rn(nx - ax.w, dx) + __fdiv_rn(ay1.w - ay0.w, dy);
    if (DIM == 3) {
        o.x += __fdiv_rn(az1.x - az0.x, dz); o.y += __fdiv_rn(az1.y - az0.y, dz);
        o.z += __fdiv_rn(az1.z - az0.z, dz); o.w += __fdiv_rn(az1.w - az0.w, dz);
    }
    float* dst = div + (long long)L.b * cf.sb + (long long)L.z * cf.sz + (long long)L.y * cf.sy + x0;
    const int nvalid = g.n[0] - x0;
    if (nvalid >= 4) *reinterpret_cast<float4*>(dst) = o;
    else for (int j = 0; j < nvalid; ++j) dst[j] = f4_get(o, j);
}

int phi_launch_divergence_vec(const DGrid& g, const DVec& v, const DField& cf, float* div, cudaStream_t s)
{
    dim3 grid((g.fext[0] / 4 + 31) / 32, (g.fext[1] + FK_WARPS - 1) / FK_WARPS, g.fext[2] * g.batch);
    if (g.dim == 3) k_div_vec<3><<<grid, FK_THREADS, 0, s>>>(g, v, cf, div);
    else            k_div_vec<2><<<grid, FK_THREADS, 0, s>>>(g, v, cf, div);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// A6  v_d[f] = vin_d[f] - (p[upper(f)] - p[lower(f)]) / dx_d on the stored faces       28 B/cell (3-D), 20 (2-D)
// Out of place (vout may equal vin): the fused step writes the projected velocity straight back into the caller's arrays.
// ---------------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(FK_THREADS)
k_gradsub_vec(DGrid g, DVec vin, DVecOut vout, DField pf, const float* __restrict__ p)
{
    const FkLine L = fk_line4<DIM>(g);
    if (!L.ok || (DIM == 3 && (L.z < 0 || L.z >= g.fext[2] - 2 * g.halo))) return;
    const int lane = threadIdx.x & 31;
    const int x0 = L.x0;
    const bool in_line = x0 < g.fext[0];
    const RowRef<3> r0 = fk_row<DIM>(g, pf, L.b, L.y, L.z);
    float4 pc = f4_splat(0.f);
    if (in_line) pc = fk_ld4(p, r0, pf, x0);
    float pl = __shfl_up_sync(0xffffffffu, pc.w, 1);                         // p[x0 - 1]
    if (in_line && (lane == 0 || x0 == 0)) pl = fk_ldx(p, r0, pf, x0 - 1);
    if (!in_line) return;
    const long long off = (long long)L.b * vin.f[0].sb + (long long)L.z * vin.f[0].sz + (long long)L.y * vin.f[0].sy + x0;
    const float dx = g.dx[0], dy = g.dx[1], dz = g.dx[2];
    const bool yz_c = L.y < g.n[1] && (DIM == 2 || L.z < g.n[2]);            // line inside the cell range in y and z
    {   // x component: faces lo..hi along x, cells along y, z
        const DField& f = vin.f[0];
        if (yz_c && x0 <= f.hi[0] && x0 + 3 >= f.lo[0]) {
            float4 a = *reinterpret_cast<const float4*>(vin.p[0] + off);
            a.x -= __fdiv_rn(pc.x - pl, dx); a.y -= __fdiv_rn(pc.y - pc.x, dx);
            a.z -= __fdiv_rn(pc.z - pc.y, dx); a.w -= __fdiv_rn(pc.w - pc.z, dx);
            if (x0 >= f.lo[0] && x0 + 3 <= f.hi[0]) *reinterpret_cast<float4*>(vout.p[0] + off) = a;
            else for (int j = 0; j < 4; ++j) if (x0 + j >= f.lo[0] && x0 + j <= f.hi[0]) vout.p[0][off + j] = f4_get(a, j);
        }
    }
    const int nvx = g.n[0] - x0;                                            // cells of this group inside the line
    if (nvx <= 0) return;
    {   // y component
        const DField& f = vin.f[1];
        if (L.y >= f.lo[1] && L.y <= f.hi[1] && (DIM == 2 || L.z < g.n[2])) {
            const RowRef<3> rm = fk_row<DIM>(g, pf, L.b, L.y - 1, L.z);
            const float4 pm = fk_ld4(p, rm, pf, x0);
            float4 a = *reinterpret_cast<const float4*>(vin.p[1] + off);
            a.x -= __fdiv_rn(pc.x - pm.x, dy); a.y -= __fdiv_rn(pc.y - pm.y, dy);
            a.z -= __fdiv_rn(pc.z - pm.z, dy); a.w -= __fdiv_rn(pc.w - pm.w, dy);
            if (nvx >= 4) *reinterpret_cast<float4*>(vout.p[1] + off) = a;
            else for (int j = 0; j < nvx; ++j) vout.p[1][off + j] = f4_get(a, j);
        }
    }
    if (DIM == 3) {   // z component
        const DField& f = vin.f[2];
        if (L.z >= f.lo[2] && L.z <= f.hi[2] && L.y < g.n[1]) {
            const RowRef<3> rm = fk_row<DIM>(g, pf, L.b, L.y, L.z - 1);
            const float4 pm = fk_ld4(p, rm, pf, x0);
            float4 a = *reinterpret_cast<const float4*>(vin.p[2] + off);
            a.x -= __fdiv_rn(pc.x - pm.x, dz); a.y -= __fdiv_rn(pc.y - pm.y, dz);
            a.z -= __fdiv_rn(pc.z - pm.z, dz); a.w -= __fdiv_rn(pc.w - pm.w, dz);
            if (nvx >= 4) *reinterpret_cast<float4*>(vout.p[2] + off) = a;
            else for (int j = 0; j < nvx; ++j) vout.p[2][off + j] = f4_get(a, j);
        }
    }
}

int phi_launch_grad_sub_vec(const DGrid& g, const DVec& vin, const DVecOut& vout, const DField& pf, const float* p, cudaStream_t s)
{
    dim3 grid((g.fext[0] / 4 + 31) / 32, (g.fext[1] + FK_WARPS - 1) / FK_WARPS, g.fext[2] * g.batch);
    if (g.dim == 3) k_gradsub_vec<3><<<grid, FK_THREADS, 0, s>>>(g, vin, vout, pf, p);
    else            k_gradsub_vec<2><<<grid, FK_THREADS, 0, s>>>(g, vin, vout, pf, p);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// A9-A11  semi-Lagrangian advection, lanes = consecutive x
// ---------------------------------------------------------------------------------------------------------
struct FkLookup { int i[3]; float t[3]; };

template <int DIM>
__device__ __forceinline__ void fk_lookup_axis(FkLookup& L, int a, int idx, float v, float dt, float dxa)
{
    const float delta = __fdiv_rn(-dt * v, dxa);          // displacement in cells (advect.py:20-24, _resample.py:257-258)
    const float fl = floorf(delta);
    L.i[a] = idx + (int)fl;
    L.t[a] = delta - fl;
}

// n-linear interpolation at the looked-up position: weights = products of frac / (1 - frac), weighted sum over the 2^d
// neighbours in the order of the scalar kernel (PhiML/phiml/math/_ops.py:1010-1014)
template <int DIM, bool LIMITS>
__device__ __forceinline__ float fk_interp(const float* __restrict__ a, const DGrid& g, const DField& f, int b, const FkLookup& L,
                                           float* vmin, float* vmax)
{
    // planes readable below / above the owned z range on a slab (PHI_BC_HALO) count as stored values
    const int zlo = f.lo[2] - ((DIM == 3 && f.klo[2] == PHI_BC_HALO) ? f.halo : 0);
    const int zhi = f.hi[2] + ((DIM == 3 && f.khi[2] == PHI_BC_HALO) ? f.halo : 0);
    bool inside = L.i[0] >= f.lo[0] && L.i[0] + 1 <= f.hi[0] && L.i[1] >= f.lo[1] && L.i[1] + 1 <= f.hi[1];
    if (DIM == 3) inside = inside && L.i[2] >= zlo && L.i[2] + 1 <= zhi;
    float acc = 0.f, mn = 3.4e38f, mx = -3.4e38f;
    const float tx = L.t[0], ty = L.t[1], tz = L.t[2];
    if (inside) {
        const float* p = a + (long long)b * f.sb + (DIM == 3 ? (long long)L.i[2] * f.sz : 0) + (long long)L.i[1] * f.sy + L.i[0];
        const float n00 = __ldg(p), n10 = __ldg(p + 1), n01 = __ldg(p + f.sy), n11 = __ldg(p + f.sy + 1);
        if (DIM == 3) {
            const float m00 = __ldg(p + f.sz), m10 = __ldg(p + f.sz + 1), m01 = __ldg(p + f.sz + f.sy), m11 = __ldg(p + f.sz + f.sy + 1);
            const float w00 = (1.f - tx) * (1.f - ty), w01 = (1.f - tx) * ty, w10 = tx * (1.f - ty), w11 = tx * ty;
            acc += n00 * (w00 * (1.f - tz)); acc += m00 * (w00 * tz);
            acc += n01 * (w01 * (1.f - tz)); acc += m01 * (w01 * tz);
            acc += n10 * (w10 * (1.f - tz)); acc += m10 * (w10 * tz);
            acc += n11 * (w11 * (1.f - tz)); acc += m11 * (w11 * tz);
            if (LIMITS) { mn = fminf(fminf(fminf(n00, n10), fminf(n01, n11)), fminf(fminf(m00, m10), fminf(m01, m11)));
                          mx = fmaxf(fmaxf(fmaxf(n00, n10), fmaxf(n01, n11)), fmaxf(fmaxf(m00, m10), fmaxf(m01, m11))); }
        } else {
            acc += n00 * ((1.f - tx) * (1.f - ty)); acc += n01 * ((1.f - tx) * ty);
            acc += n10 * (tx * (1.f - ty)); acc += n11 * (tx * ty);
            if (LIMITS) { mn = fminf(fminf(n00, n10), fminf(n01, n11)); mx = fmaxf(fmaxf(n00, n10), fmaxf(n01, n11)); }
        }
    } else {
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
            const float wx = cx ? tx : 1.f - tx;
#pragma unroll
            for (int cy = 0; cy < 2; ++cy) {
                const float wxy = wx * (cy ? ty : 1.f - ty);
                if (DIM == 3) {
#pragma unroll
                    for (int cz = 0; cz < 2; ++cz) {
                        const float w = wxy * (cz ? tz : 1.f - tz);
                        const float n = phi_fetch<DIM>(a, g, f, b, L.i[0] + cx, L.i[1] + cy, L.i[2] + cz);
                        acc += n * w;
                        if (LIMITS) { mn = fminf(mn, n); mx = fmaxf(mx, n); }
                    }
                } else {
                    const float n = phi_fetch<DIM>(a, g, f, b, L.i[0] + cx, L.i[1] + cy, 0);
                    acc += n * wxy;
                    if (LIMITS) { mn = fminf(mn, n); mx = fmaxf(mx, n); }
                }
            }
        }
    }
    if (LIMITS) { *vmin = mn; *vmax = mx; }
    return acc;
}

#define FK_XCHUNKS 4        // a warp walks 4 chunks of 32 cells of its line

struct FkAdvLine { int b, y, z, xs; bool ok; };

template <int DIM>
__device__ __forceinline__ FkAdvLine fk_adv_line(const DGrid& g)
{
    FkAdvLine L;
    L.xs = blockIdx.x * 32 * FK_XCHUNKS;
    L.y = blockIdx.y * FK_WARPS + (threadIdx.x >> 5);
    const int zb = blockIdx.z;
    if (DIM == 3) { L.z = zb % g.fext[2] - g.halo; L.b = zb / g.fext[2]; } else { L.z = 0; L.b = zb; }
    L.ok = L.y < g.fext[1] && (DIM == 2 || (L.z >= 0 && L.z < g.fext[2] - 2 * g.halo));
    return L;
}

// x-shifted copies of a per-lane line value: v(x + 1) / v(x - 1).  Edge lanes of the 32-cell chunk load instead.
__device__ __forceinline__ float fk_next(float own, const float* __restrict__ a, const RowRef<3>& r, const DField& f, int x)
{
    float n = __shfl_down_sync(0xffffffffu, own, 1);
    if ((threadIdx.x & 31) == 31) n = fk_ldx(a, r, f, x + 1);
    return n;
}
__device__ __forceinline__ float fk_prev(float own, const float* __restrict__ a, const RowRef<3>& r, const DField& f, int x)
{
    float n = __shfl_up_sync(0xffffffffu, own, 1);
    if ((threadIdx.x & 31) == 0) n = fk_ldx(a, r, f, x - 1);
    return n;
}

__device__ __forceinline__ float fk_avg4(float f00, float f10, float f01, float f11, bool a_first)
{
    // shift resampling = sample_subgrid lerps the axes in spatial order (PhiML/phiml/math/_nd.py:973-1003):
    // f{da}{dt}: da in {0,1} along the component's own axis a, dt in {-1 -> 0, 0 -> 1} along the target face axis t
    if (a_first) { const float u0 = f10 * 0.5f + f00 * 0.5f, u1 = f11 * 0.5f + f01 * 0.5f; return u1 * 0.5f + u0 * 0.5f; }
    const float w0 = f01 * 0.5f + f00 * 0.5f, w1 = f11 * 0.5f + f10 * 0.5f;
    return w1 * 0.5f + w0 * 0.5f;
}

// Centred field: dst = interp(src, x - dt v(x)) [+ add_scale * add]      (smoke advection + inflow of the notebook step)
template <int DIM>
__global__ void __launch_bounds__(FK_THREADS)
k_advect_centered_vec(DGrid g, DVec vel, DField ff, const float* __restrict__ src, float* __restrict__ dst, float dt,
                      const float* __restrict__ add, float add_scale)
{
    const FkAdvLine L = fk_adv_line<DIM>(g);
    if (!L.ok || L.y >= g.n[1] || (DIM == 3 && L.z >= g.n[2])) return;
    const int lane = threadIdx.x & 31;
    const RowRef<3> rx = fk_row<DIM>(g, vel.f[0], L.b, L.y, L.z);
    const RowRef<3> ry0 = fk_row<DIM>(g, vel.f[1], L.b, L.y, L.z), ry1 = fk_row<DIM>(g, vel.f[1], L.b, L.y + 1, L.z);
    RowRef<3> rz0 = rx, rz1 = rx;
    if (DIM == 3) { rz0 = fk_row<DIM>(g, vel.f[2], L.b, L.y, L.z); rz1 = fk_row<DIM>(g, vel.f[2], L.b, L.y, L.z + 1); }
    const long long line = (long long)L.b * ff.sb + (long long)L.z * ff.sz + (long long)L.y * ff.sy;
#pragma unroll 2
    for (int j = 0; j < FK_XCHUNKS; ++j) {
        const int xb = L.xs + 32 * j;
        if (xb >= g.n[0]) break;                                   // warp-uniform
        const int x = xb + lane;
        const float a0 = fk_ldx(vel.p[0], rx, vel.f[0], x);
        const float a0p = fk_next(a0, vel.p[0], rx, vel.f[0], x);
        const float b0 = fk_ldx(vel.p[1], ry0, vel.f[1], x), b1 = fk_ldx(vel.p[1], ry1, vel.f[1], x);
        FkLookup K; K.i[2] = 0; K.t[2] = 0.f;
        fk_lookup_axis<DIM>(K, 0, x, a0p * 0.5f + a0 * 0.5f, dt, g.dx[0]);
        fk_lookup_axis<DIM>(K, 1, L.y, b1 * 0.5f + b0 * 0.5f, dt, g.dx[1]);
        if (DIM == 3) {
            const float c0 = fk_ldx(vel.p[2], rz0, vel.f[2], x), c1 = fk_ldx(vel.p[2], rz1, vel.f[2], x);
            fk_lookup_axis<DIM>(K, 2, L.z, c1 * 0.5f + c0 * 0.5f, dt, g.dx[2]);
        }
        if (x >= g.n[0]) continue;
        float r = fk_interp<DIM, false>(src, g, ff, L.b, K, nullptr, nullptr);
        if (add) r = r + add_scale * __ldg(add + line + x);
        dst[line + x] = r;
    }
}

// Staggered field, all components in one launch: dst_c = interp(src_c, face_c - dt v(face_c)) [+ dt * buoyancy_c]
//   buoyancy_c = (s * b_c)[upper cell] * 0.5 + (s * b_c)[lower cell] * 0.5      (sample_grid_at_faces)
template <int DIM, bool BUOY>
__global__ void __launch_bounds__(FK_THREADS)
k_advect_staggered_vec(DGrid g, DVec vel, DVec fld, DVecOut dst, float dt, DField sf, const float* __restrict__ s,
                       float b0, float b1, float b2)
{
    const FkAdvLine L = fk_adv_line<DIM>(g);
    if (!L.ok) return;
    const int lane = threadIdx.x & 31;
    const int y = L.y, z = L.z, b = L.b;
    const float* vx = vel.p[0]; const float* vy = vel.p[1]; const float* vz = vel.p[2];
    const DField& fx = vel.f[0]; const DField& fy = vel.f[1]; const DField& fz = vel.f[2];
    // the 11 (3-D) / 4 (2-D) velocity lines the three faces of a cell read
    const RowRef<3> rA0 = fk_row<DIM>(g, fx, b, y, z), rA1 = fk_row<DIM>(g, fx, b, y - 1, z);
    const RowRef<3> rB0 = fk_row<DIM>(g, fy, b, y, z), rB1 = fk_row<DIM>(g, fy, b, y + 1, z);
    RowRef<3> rA2 = rA0, rB2 = rB0, rB3 = rB0, rC0 = rA0, rC1 = rA0, rC2 = rA0, rC3 = rA0;
    if (DIM == 3) {
        rA2 = fk_row<DIM>(g, fx, b, y, z - 1);
        rB2 = fk_row<DIM>(g, fy, b, y, z - 1); rB3 = fk_row<DIM>(g, fy, b, y + 1, z - 1);
        rC0 = fk_row<DIM>(g, fz, b, y, z); rC1 = fk_row<DIM>(g, fz, b, y, z + 1);
        rC2 = fk_row<DIM>(g, fz, b, y - 1, z); rC3 = fk_row<DIM>(g, fz, b, y - 1, z + 1);
    }
    // which components store a face on this line (warp-uniform)
    const bool on0 = y >= fld.f[0].lo[1] && y <= fld.f[0].hi[1] && (DIM == 2 || (z >= fld.f[0].lo[2] && z <= fld.f[0].hi[2]));
    const bool on1 = y >= fld.f[1].lo[1] && y <= fld.f[1].hi[1] && (DIM == 2 || (z >= fld.f[1].lo[2] && z <= fld.f[1].hi[2]));
    const bool on2 = DIM == 3 && y >= fld.f[2].lo[1] && y <= fld.f[2].hi[1] && z >= fld.f[2].lo[2] && z <= fld.f[2].hi[2];
    if (!(on0 || on1 || on2)) return;
    // buoyancy lines of the centred field (upper cell = this index, lower cell = index - e_c)
    RowRef<3> sC = rA0, sY = rA0, sZ = rA0;
    if (BUOY) {
        sC = fk_row<DIM>(g, sf, b, y, z);
        if (b1 != 0.f) sY = fk_row<DIM>(g, sf, b, y - 1, z);
        if (DIM == 3 && b2 != 0.f) sZ = fk_row<DIM>(g, sf, b, y, z - 1);
    }
    const long long line = (long long)b * fx.sb + (long long)z * fx.sz + (long long)y * fx.sy;
    const int xend = g.fext[0];
#pragma unroll 1
    for (int j = 0; j < FK_XCHUNKS; ++j) {
        const int xb = L.xs + 32 * j;
        if (xb >= xend) break;                                     // warp-uniform
        const int x = xb + lane;
        const float A0 = fk_ldx(vx, rA0, fx, x), A1 = fk_ldx(vx, rA1, fx, x);
        const float B0 = fk_ldx(vy, rB0, fy, x), B1 = fk_ldx(vy, rB1, fy, x);
        const float A0p = fk_next(A0, vx, rA0, fx, x), A1p = fk_next(A1, vx, rA1, fx, x);
        const float B0m = fk_prev(B0, vy, rB0, fy, x), B1m = fk_prev(B1, vy, rB1, fy, x);
        float A2 = 0.f, A2p = 0.f, B2 = 0.f, B3 = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, C3 = 0.f, C0m = 0.f, C1m = 0.f;
        if (DIM == 3) {
            A2 = fk_ldx(vx, rA2, fx, x); B2 = fk_ldx(vy, rB2, fy, x); B3 = fk_ldx(vy, rB3, fy, x);
            C0 = fk_ldx(vz, rC0, fz, x); C1 = fk_ldx(vz, rC1, fz, x); C2 = fk_ldx(vz, rC2, fz, x); C3 = fk_ldx(vz, rC3, fz, x);
            A2p = fk_next(A2, vx, rA2, fx, x);
            C0m = fk_prev(C0, vz, rC0, fz, x); C1m = fk_prev(C1, vz, rC1, fz, x);
        }
        float sc = 0.f, scm = 0.f;
        if (BUOY) {
            sc = fk_ldx(s, sC, sf, x);
            if (b0 != 0.f) scm = fk_prev(sc, s, sC, sf, x);
        }
        if (x >= xend) continue;
        FkLookup K; K.i[2] = 0; K.t[2] = 0.f;
        if (on0 && x >= fld.f[0].lo[0] && x <= fld.f[0].hi[0]) {          // x faces: target axis t = 0
            fk_lookup_axis<DIM>(K, 0, x, A0, dt, g.dx[0]);
            fk_lookup_axis<DIM>(K, 1, y, fk_avg4(B0m, B1m, B0, B1, false), dt, g.dx[1]);
            if (DIM == 3) fk_lookup_axis<DIM>(K, 2, z, fk_avg4(C0m, C1m, C0, C1, false), dt, g.dx[2]);
            float r = fk_interp<DIM, false>(fld.p[0], g, fld.f[0], b, K, nullptr, nullptr);
            if (BUOY && b0 != 0.f) r = r + ((sc * b0) * 0.5f + (scm * b0) * 0.5f) * dt;
            dst.p[0][line + x] = r;
        }
        if (on1 && x >= fld.f[1].lo[0] && x <= fld.f[1].hi[0]) {          // y faces: t = 1
            fk_lookup_axis<DIM>(K, 0, x, fk_avg4(A1, A1p, A0, A0p, true), dt, g.dx[0]);
            fk_lookup_axis<DIM>(K, 1, y, B0, dt, g.dx[1]);
            if (DIM == 3) fk_lookup_axis<DIM>(K, 2, z, fk_avg4(C2, C3, C0, C1, false), dt, g.dx[2]);
            float r = fk_interp<DIM, false>(fld.p[1], g, fld.f[1], b, K, nullptr, nullptr);
            if (BUOY && b1 != 0.f) r = r + ((sc * b1) * 0.5f + (fk_ldx(s, sY, sf, x) * b1) * 0.5f) * dt;
            dst.p[1][line + x] = r;
        }
        if (on2 && x >= fld.f[2].lo[0] && x <= fld.f[2].hi[0]) {          // z faces: t = 2
            fk_lookup_axis<DIM>(K, 0, x, fk_avg4(A2, A2p, A0, A0p, true), dt, g.dx[0]);
            fk_lookup_axis<DIM>(K, 1, y, fk_avg4(B2, B3, B0, B1, true), dt, g.dx[1]);
            fk_lookup_axis<DIM>(K, 2, z, C0, dt, g.dx[2]);
            float r = fk_interp<DIM, false>(fld.p[2], g, fld.f[2], b, K, nullptr, nullptr);
            if (BUOY && b2 != 0.f) r = r + ((sc * b2) * 0.5f + (fk_ldx(s, sZ, sf, x) * b2) * 0.5f) * dt;
            dst.p[2][line + x] = r;
        }
    }
}

static dim3 adv_grid(const DGrid& g)
{
    return dim3((g.fext[0] + 32 * FK_XCHUNKS - 1) / (32 * FK_XCHUNKS), (g.fext[1] + FK_WARPS - 1) / FK_WARPS, g.fext[2] * g.batch);
}

int phi_launch_advect_centered_vec(const DGrid& g, const DVec& vel, const DField& ff, const float* src, float* dst, float dt,
                                   const float* add, float add_scale, cudaStream_t s)
{
    if (g.dim == 3) k_advect_centered_vec<3><<<adv_grid(g), FK_THREADS, 0, s>>>(g, vel, ff, src, dst, dt, add, add_scale);
    else            k_advect_centered_vec<2><<<adv_grid(g), FK_THREADS, 0, s>>>(g, vel, ff, src, dst, dt, add, add_scale);
    return (int)cudaGetLastError();
}

int phi_launch_advect_staggered_vec(const DGrid& g, const DVec& vel, const DVec& fld, const DVecOut& dst, float dt,
                                    const DField* sf, const float* sarr, const float bu[3], cudaStream_t s)
{
    const bool buoy = sarr != nullptr && bu && (bu[0] != 0.f || bu[1] != 0.f || (g.dim == 3 && bu[2] != 0.f));
    const DField sfv = sf ? *sf : fld.f[0];
    const float b0 = buoy ? bu[0] : 0.f, b1 = buoy ? bu[1] : 0.f, b2 = (buoy && g.dim == 3) ? bu[2] : 0.f;
    if (g.dim == 3) {
        if (buoy) k_advect_staggered_vec<3, true><<<adv_grid(g), FK_THREADS, 0, s>>>(g, vel, fld, dst, dt, sfv, sarr, b0, b1, b2);
        else      k_advect_staggered_vec<3, false><<<adv_grid(g), FK_THREADS, 0, s>>>(g, vel, fld, dst, dt, sfv, sarr, b0, b1, b2);
    } else {
        if (buoy) k_advect_staggered_vec<2, true><<<adv_grid(g), FK_THREADS, 0, s>>>(g, vel, fld, dst, dt, sfv, sarr, b0, b1, b2);
        else      k_advect_staggered_vec<2, false><<<adv_grid(g), FK_THREADS, 0, s>>>(g, vel, fld, dst, dt, sfv, sarr, b0, b1, b2);
    }
    return (int)cudaGetLastError();
}
