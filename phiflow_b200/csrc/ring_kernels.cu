// TMA-staged shared-memory ring for the 5/7-point stencil: the B200 fast path of laplace (A7) and of the CG solve (A12).
//
// Why: the stencil kernels are pure HBM streams (8..20 B/cell).  Saturating 6.5 TB/s needs ~50 KB in flight per SM; with
// register-staged loads that costs occupancy the CG passes do not have.  Here ONE producer warp per CTA streams whole
// grid lines into a ring of shared-memory stages with cp.async.bulk (the TMA engine; SASS: UBLKCP) completing on
// mbarriers, and eight consumer warps compute out of shared memory.  Loads never occupy registers, the pipeline depth
// is a launch parameter, and every byte is fetched from L2/HBM exactly once per tile (+ halo lines).
//
// Tile = TY consecutive grid lines (full x extent) of one z plane (3-D) or of one image (2-D).  In 3-D a CTA marches a
// chunk of planes; plane z needs planes z-1, z, z+1, which sit in three consecutive ring slots.  Halo lines y0-1 / y0+TY
// and halo planes are fetched from wherever the boundary condition says the ghost values live:
//   PERIODIC -> the wrapped line, ZERO_GRADIENT -> the clamped line, constant -> not fetched (consumers substitute c).
// x ghosts are read from the staged line itself.  Outputs go straight from registers to global memory (STG.128).
#include <cooperative_groups.h>
#include "cg_common.cuh"
#include "launch.cuh"

namespace cg = cooperative_groups;

#define RING_CONSUMERS 256
#define RING_THREADS (RING_CONSUMERS + 32)
#define RING_MAX_STAGES 8

struct RingCfg {
    int TY, R, pitch, nx4;
    int stage_floats;          // floats between consecutive stages
    int ZC, nyt, nzc;
    int units_per_batch, total_units;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- ring state (per thread, registers) ------------------------------------------------------------------------------
struct Ring {
    float* stage0;
    uint32_t stage0_s, full0, empty0;
    unsigned idx;              // planes produced (producer) / fully processed base index (consumers)
};

__device__ __forceinline__ void ring_init(Ring& rg, unsigned char* smem, const RingCfg& cfg)
{
    // layout: [R full barriers][R empty barriers][pad to 128][stages]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    rg.full0 = smem_u32(bars);
    rg.empty0 = smem_u32(bars + RING_MAX_STAGES);
    rg.stage0 = reinterpret_cast<float*>(smem + 128);
    rg.stage0_s = smem_u32(rg.stage0);
    rg.idx = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg.R; ++s) {
            mbar_init(rg.full0 + 8 * s, 1);
            mbar_init(rg.empty0 + 8 * s, RING_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// Producer warp: stage the lines of one plane.  hsrc: NH haloed arrays (TY+2 lines), esrc: NE element-wise arrays (TY lines).
template <int DIM>
__device__ __forceinline__ void ring_produce(Ring& rg, const RingCfg& cfg, const DGrid& g, const DField& pf,
                                             int NH, int NE, const float* const* hsrc, const float* const* esrc,
                                             int b, int y0, int z, bool interior)
{
    const int lane = threadIdx.x & 31;
    const unsigned slot = rg.idx % cfg.R, use = rg.idx / cfg.R;
    const uint32_t full = rg.full0 + 8 * slot, empty = rg.empty0 + 8 * slot;
    const uint32_t row_bytes = (uint32_t)cfg.pitch * 4u;
    const int hrows = cfg.TY + 2;
    const int total = NH * hrows + (interior ? NE * cfg.TY : 0);
    if (lane == 0) mbar_wait(empty, (use & 1u) ^ 1u);
    __syncwarp();
    const float* src[4]; uint32_t dst[4]; int cnt = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        src[it] = nullptr; dst[it] = 0;
        const int r = lane + 32 * it;
        if (r < total) {
            int arr, j, yy; const float* base;
            if (r < NH * hrows) { arr = r / hrows; j = r - arr * hrows; yy = y0 - 1 + j; base = hsrc[arr];
                                  if (yy <= g.n[1]) {
                                      RowRef<DIM> row = DIM == 3 ? phi_row<DIM>(g, pf, b, yy, z) : phi_row<DIM>(g, pf, b, yy, 0);
                                      if (row.off >= 0) { src[it] = base + row.off; dst[it] = rg.stage0_s + 4u * (slot * cfg.stage_floats + (arr * hrows + j) * cfg.pitch); }
                                  } }
            else { const int q = r - NH * hrows; arr = q / cfg.TY; j = q - arr * cfg.TY; yy = y0 + j; base = esrc[arr];
                   if (yy < g.n[1]) {
                       const long long off = (long long)b * pf.sb + (DIM == 3 ? (long long)z * pf.sz : 0) + (long long)yy * pf.sy;
                       src[it] = base + off; dst[it] = rg.stage0_s + 4u * (slot * cfg.stage_floats + (NH * hrows + arr * cfg.TY + j) * cfg.pitch);
                   } }
        }
        cnt += src[it] != nullptr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) { if (cnt > 0) mbar_expect_tx(full, (uint32_t)cnt * row_bytes); else mbar_arrive(full); }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it)
        if (src[it]) bulk_g2s(dst[it], src[it], row_bytes, full);
    rg.idx++;
}

__device__ __forceinline__ void ring_wait_full(const Ring& rg, const RingCfg& cfg, unsigned idx)
{ mbar_wait(rg.full0 + 8 * (idx % cfg.R), (idx / cfg.R) & 1u); }

__device__ __forceinline__ void ring_release(const Ring& rg, const RingCfg& cfg, unsigned idx)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(rg.empty0 + 8 * (idx % cfg.R));
}

__device__ __forceinline__ const float* ring_slot(const Ring& rg, const RingCfg& cfg, unsigned idx)
{ return rg.stage0 + (size_t)(idx % cfg.R) * cfg.stage_floats; }

// ---- consumer: one plane of one tile ---------------------------------------------------------------------------------
// sm/sc/sp: slots holding planes z-1, z, z+1 (sm/sp unused in 2-D).  Values of the differenced array are h0 (+ beta*h1).
template <int DIM, int NH, int NE, class Epi>
__device__ __forceinline__ void ring_compute(const RingCfg& cfg, const DGrid& g, const DField& pf,
                                             const float* sm, const float* sc, const float* sp, float beta,
                                             int b, int y0, int z, Epi& epi)
{
    const int pitch = cfg.pitch, hrows = cfg.TY + 2;
    const int total = cfg.TY * cfg.nx4;
    const int nx = g.n[0], ny = g.n[1];
    const int h1 = hrows * pitch;                        // offset of the second haloed array inside a stage
    const int e0off = NH * hrows * pitch;
    const float ix2 = g.inv_dx2[0], iy2 = g.inv_dx2[1], iz2 = g.inv_dx2[2];
    const bool zm_const = DIM == 3 && z - 1 < 0 && pf.klo[2] == PHI_BC_CONST;
    const bool zp_const = DIM == 3 && z + 1 > g.n[2] - 1 && pf.khi[2] == PHI_BC_CONST;
    const bool use1 = NH == 2 && beta != 0.f;

    auto val4 = [&](const float* s, int off) -> float4 {
        float4 a = *reinterpret_cast<const float4*>(s + off);
        if (use1) { const float4 o = *reinterpret_cast<const float4*>(s + h1 + off);
                    a.x += beta * o.x; a.y += beta * o.y; a.z += beta * o.z; a.w += beta * o.w; }
        return a;
    };
    auto val1 = [&](const float* s, int off) -> float {
        float a = s[off];
        if (use1) a += beta * s[h1 + off];
        return a;
    };

    for (int gi = threadIdx.x; gi < total; gi += RING_CONSUMERS) {
        const int j = gi / cfg.nx4, x4 = gi - j * cfg.nx4;
        const int y = y0 + j, x0 = x4 * 4;
        if (y >= ny || x0 >= nx) continue;
        const int nvalid = min(4, nx - x0);
        const int row = (j + 1) * pitch;
        const int rc = row + x0;
        const float4 c = val4(sc, rc);
        const float4 ym = (y - 1 < 0 && pf.klo[1] == PHI_BC_CONST) ? f4_splat(pf.clo[1]) : val4(sc, rc - pitch);
        const float4 yp = (y + 1 > ny - 1 && pf.khi[1] == PHI_BC_CONST) ? f4_splat(pf.chi[1]) : val4(sc, rc + pitch);
        float xl, xr;
        if (x0 > 0) xl = val1(sc, rc - 1);
        else { const int k = pf.klo[0]; xl = k == PHI_BC_PERIODIC ? val1(sc, row + nx - 1) : (k == PHI_BC_ZERO_GRADIENT ? c.x : pf.clo[0]); }
        if (x0 + 4 < nx) xr = val1(sc, rc + 4);
        else { const int k = pf.khi[0]; xr = k == PHI_BC_PERIODIC ? val1(sc, row) : (k == PHI_BC_ZERO_GRADIENT ? f4_get(c, nvalid - 1) : pf.chi[0]); }
        float4 l4 = make_float4(xl, c.x, c.y, c.z);
        float4 r4 = make_float4(c.y, c.z, c.w, xr);
        if (nvalid < 4) f4_set(r4, nvalid - 1, xr);
        float4 q;
        q.x = (l4.x + r4.x - 2.f * c.x) * ix2 + (ym.x + yp.x - 2.f * c.x) * iy2;
        q.y = (l4.y + r4.y - 2.f * c.y) * ix2 + (ym.y + yp.y - 2.f * c.y) * iy2;
        q.z = (l4.z + r4.z - 2.f * c.z) * ix2 + (ym.z + yp.z - 2.f * c.z) * iy2;
        q.w = (l4.w + r4.w - 2.f * c.w) * ix2 + (ym.w + yp.w - 2.f * c.w) * iy2;
        if (DIM == 3) {
            const float4 zm = zm_const ? f4_splat(pf.clo[2]) : val4(sm, rc);
            const float4 zp = zp_const ? f4_splat(pf.chi[2]) : val4(sp, rc);
            q.x += (zm.x + zp.x - 2.f * c.x) * iz2;
            q.y += (zm.y + zp.y - 2.f * c.y) * iz2;
            q.z += (zm.z + zp.z - 2.f * c.z) * iz2;
            q.w += (zm.w + zp.w - 2.f * c.w) * iz2;
        }
        float4 e0 = f4_splat(0.f), e1 = f4_splat(0.f);
        if (NE >= 1) e0 = *reinterpret_cast<const float4*>(sc + e0off + j * pitch + x0);
        if (NE >= 2) e1 = *reinterpret_cast<const float4*>(sc + e0off + (cfg.TY + j) * pitch + x0);
        const long long off = (long long)b * pf.sb + (DIM == 3 ? (long long)z * pf.sz : 0) + (long long)y * pf.sy + x0;
        epi(off, c, q, nvalid, e0, e1);
    }
}

// ---- one unit: producer streams its planes, consumers march through them --------------------------------------------------
struct RingUnit { int b, y0, z0, z1; };

template <int DIM>
__device__ __forceinline__ RingUnit ring_unit(const RingCfg& cfg, const DGrid& g, int unit)
{
    RingUnit u;
    u.b = unit / cfg.units_per_batch;
    const int r = unit - u.b * cfg.units_per_batch;
    if (DIM == 3) { const int yt = r % cfg.nyt, zc = r / cfg.nyt; u.y0 = yt * cfg.TY; u.z0 = zc * cfg.ZC; u.z1 = min(g.n[2], u.z0 + cfg.ZC); }
    else { u.y0 = r * cfg.TY; u.z0 = 0; u.z1 = 1; }
    return u;
}

template <int DIM, int NH, int NE, class Epi>
__device__ __forceinline__ void ring_process_unit(Ring& rg, const RingCfg& cfg, const DGrid& g, const DField& pf,
                                                  const float* const* hsrc, const float* const* esrc, float beta,
                                                  const RingUnit& u, Epi& epi)
{
    const bool producer = threadIdx.x >= RING_CONSUMERS;
    const int nh = (NH == 2 && beta == 0.f) ? 1 : NH;              // first CG iteration: d' = r, old direction not read
    if (DIM == 3) {
        const int nz = u.z1 - u.z0;
        if (producer) {
            for (int p = 0; p < nz + 2; ++p)
                ring_produce<DIM>(rg, cfg, g, pf, nh, NE, hsrc, esrc, u.b, u.y0, u.z0 - 1 + p, p >= 1 && p <= nz);
        } else {
            const unsigned base = rg.idx;
            ring_wait_full(rg, cfg, base); ring_wait_full(rg, cfg, base + 1);
            for (int zi = 0; zi < nz; ++zi) {
                ring_wait_full(rg, cfg, base + zi + 2);
                ring_compute<DIM, NH, NE>(cfg, g, pf, ring_slot(rg, cfg, base + zi), ring_slot(rg, cfg, base + zi + 1),
                                          ring_slot(rg, cfg, base + zi + 2), beta, u.b, u.y0, u.z0 + zi, epi);
                ring_release(rg, cfg, base + zi);
            }
            ring_release(rg, cfg, base + nz); ring_release(rg, cfg, base + nz + 1);
            rg.idx = base + nz + 2;
        }
    } else {
        if (producer) ring_produce<DIM>(rg, cfg, g, pf, nh, NE, hsrc, esrc, u.b, u.y0, 0, true);
        else {
            const unsigned base = rg.idx;
            ring_wait_full(rg, cfg, base);
            const float* sc = ring_slot(rg, cfg, base);
            ring_compute<DIM, NH, NE>(cfg, g, pf, sc, sc, sc, beta, u.b, u.y0, 0, epi);
            ring_release(rg, cfg, base);
            rg.idx = base + 1;
        }
    }
}

// ---- epilogues (values of element-wise arrays arrive from shared memory) -------------------------------------------------
template <bool AXPY>
struct REpiLaplace {
    float* y; float coeff;
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4&, const float4&)
    {
        float4 o = q;
        if (AXPY) { o.x = c.x + coeff * q.x; o.y = c.y + coeff * q.y; o.z = c.z + coeff * q.z; o.w = c.w + coeff * q.w; }
        if (nvalid == 4) *reinterpret_cast<float4*>(y + off) = o;
        else for (int j = 0; j < nvalid; ++j) y[off + j] = f4_get(o, j);
    }
};

struct REpiResidual0 {          // e0 = rhs
    float* r; float mean, offs; float acc0, acc1;
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4& y, const float4&)
    {
        float4 rt = make_float4((y.x - mean) - q.x, (y.y - mean) - q.y, (y.z - mean) - q.z, (y.w - mean) - q.w);
        float4 rr = make_float4(rt.x - offs, rt.y - offs, rt.z - offs, rt.w - offs);
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(r + off) = rr;
            acc0 += rr.x * rr.x + rr.y * rr.y + rr.z * rr.z + rr.w * rr.w;
            acc1 += rt.x * rt.x + rt.y * rt.y + rt.z * rt.z + rt.w * rt.w;
        } else for (int j = 0; j < nvalid; ++j) { const float a = f4_get(rr, j), t = f4_get(rt, j); r[off + j] = a; acc0 += a * a; acc1 += t * t; }
    }
};

struct REpiPassA {
    float* dnew; float acc0, acc1;
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4&, const float4&)
    {
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(dnew + off) = c;
            acc0 += c.x * q.x + c.y * q.y + c.z * q.z + c.w * q.w;
            acc1 += (c.x + c.y) + (c.z + c.w);
        } else for (int j = 0; j < nvalid; ++j) { const float v = f4_get(c, j); dnew[off + j] = v; acc0 += v * f4_get(q, j); acc1 += v; }
    }
};

struct REpiPassB {              // e0 = x, e1 = r
    float* x; float* r; float alpha, offs; float acc0, acc1;
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4& xe, const float4& re)
    {
        float4 xv = xe, rv = re;
        xv.x += alpha * c.x; xv.y += alpha * c.y; xv.z += alpha * c.z; xv.w += alpha * c.w;
        rv.x -= alpha * (q.x + offs); rv.y -= alpha * (q.y + offs); rv.z -= alpha * (q.z + offs); rv.w -= alpha * (q.w + offs);
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(x + off) = xv;
            *reinterpret_cast<float4*>(r + off) = rv;
            acc0 += rv.x * rv.x + rv.y * rv.y + rv.z * rv.z + rv.w * rv.w;
        } else for (int j = 0; j < nvalid; ++j) { x[off + j] = f4_get(xv, j); const float t = f4_get(rv, j); r[off + j] = t; acc0 += t * t; }
    }
};

// ---- laplace -----------------------------------------------------------------------------------------------------------
template <int DIM, bool AXPY>
__global__ void __launch_bounds__(RING_THREADS, 1)
k_laplace_ring(DGrid g, DField f, RingCfg cfg, const float* __restrict__ x, float* __restrict__ y, float coeff)
{
    extern __shared__ __align__(128) unsigned char smem[];
    Ring rg;
    ring_init(rg, smem, cfg);
    const float* hsrc[2] = {x, nullptr};
    const float* esrc[2] = {nullptr, nullptr};
    REpiLaplace<AXPY> epi{y, coeff};
    for (int unit = blockIdx.x; unit < cfg.total_units; unit += gridDim.x) {
        const RingUnit u = ring_unit<DIM>(cfg, g, unit);
        ring_process_unit<DIM, 1, 0>(rg, cfg, g, f, hsrc, esrc, 0.f, u, epi);
    }
}

// ---- CG ----------------------------------------------------------------------------------------------------------------
struct CgRingArgs {
    CgArgs a;
    RingCfg cfg;
    int ring_smem_offset;        // byte offset of the ring inside dynamic shared memory (after the CgShared block)
};

template <int DIM, class F>
__device__ __forceinline__ void ring_unit_cells(const RingCfg& cfg, const DGrid& g, const DField& pf, const RingUnit& u, F&& fn)
{
    // plain element-wise traversal of a unit by the consumer threads (sums, mean removal); no staging
    if (threadIdx.x >= RING_CONSUMERS) return;
    const int total = cfg.TY * cfg.nx4;
    for (int z = u.z0; z < u.z1; ++z)
        for (int gi = threadIdx.x; gi < total; gi += RING_CONSUMERS) {
            const int j = gi / cfg.nx4, x0 = (gi - j * cfg.nx4) * 4, y = u.y0 + j;
            if (y >= g.n[1] || x0 >= g.n[0]) continue;
            const long long off = (long long)u.b * pf.sb + (DIM == 3 ? (long long)z * pf.sz : 0) + (long long)y * pf.sy + x0;
            fn(off, min(4, g.n[0] - x0));
        }
}

template <int DIM>
__global__ void __launch_bounds__(RING_THREADS, 1)
k_cg_ring(CgRingArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const CgArgs& a = A.a;
    const RingCfg& cfg = A.cfg;
    CgShared sh = cg_carve(smem_raw, a.g.batch);
    Ring rg;
    ring_init(rg, smem_raw + A.ring_smem_offset, cfg);
    cg::grid_group grid = cg::this_grid();
    const DGrid& g = a.g;
    const int batch = g.batch;
    const double cells = (double)g.n[0] * g.n[1] * g.n[2];
    const float coffs = a.prm.matrix_offset;
    int region = 0;

    auto sweep = [&](const unsigned char* active, auto&& body) {
        int cur_b = -1; float acc0 = 0.f, acc1 = 0.f;
        for (int unit = blockIdx.x; unit < cfg.total_units; unit += gridDim.x) {
            const RingUnit u = ring_unit<DIM>(cfg, g, unit);
            if (active && !active[u.b]) continue;
            if (u.b != cur_b) {
                if (cur_b >= 0) flush_partials(sh, a.partials, region, batch, cur_b, acc0, acc1);
                cur_b = u.b; acc0 = 0.f; acc1 = 0.f;
            }
            body(u, acc0, acc1);
        }
        if (cur_b >= 0) flush_partials(sh, a.partials, region, batch, cur_b, acc0, acc1);
    };
    auto barrier_and_reduce = [&](const unsigned char* active) {
        fence_proxy_async();                       // generic-proxy stores of this pass -> later TMA (async proxy) loads
        grid.sync();
        fence_proxy_async();
        reduce_partials(sh, a.partials, region, batch, cfg.units_per_batch, active);
        region ^= 1;
    };

    for (int b = threadIdx.x; b < batch; b += blockDim.x) { sh.mean[b] = 0.f; sh.offs[b] = 0.f; }
    __syncthreads();

    if (a.prm.balance_rhs || coffs != 0.f) {
        sweep(nullptr, [&](const RingUnit& u, float& acc0, float& acc1) {
            ring_unit_cells<DIM>(cfg, g, a.pf, u, [&](long long off, int nvalid) {
                for (int j = 0; j < nvalid; ++j) { acc0 += a.rhs[off + j]; acc1 += a.x[off + j]; }
            });
        });
        barrier_and_reduce(nullptr);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            sh.mean[b] = a.prm.balance_rhs ? (float)(sh.sum0[b] / cells) : 0.f;
            sh.offs[b] = coffs * (float)sh.sum1[b];
        }
        __syncthreads();
    }

    {   // r0 = y - (A + c 11^T) x0
        const float* hsrc[2] = {a.x, nullptr};
        const float* esrc[2] = {a.rhs, nullptr};
        sweep(nullptr, [&](const RingUnit& u, float& acc0, float& acc1) {
            REpiResidual0 epi{a.r, sh.mean[u.b], sh.offs[u.b], 0.f, 0.f};
            ring_process_unit<DIM, 1, 1>(rg, cfg, g, a.pf, hsrc, esrc, 0.f, u, epi);
            acc0 += epi.acc0; acc1 += epi.acc1;
        });
    }
    barrier_and_reduce(nullptr);
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        const double d0 = sh.sum0[b], d0tol = sh.sum1[b];
        sh.delta[b] = d0;
        const float tol = fmaxf(a.prm.rtol * a.prm.rtol * (float)d0tol, a.prm.atol * a.prm.atol);
        sh.tol_sq[b] = tol; sh.rsq0[b] = (float)d0;
        const bool conv = (float)d0 <= tol;
        const bool divg = !isfinite((float)d0);
        sh.conv[b] = conv; sh.divg[b] = divg; sh.iters[b] = 0;
        sh.cont[b] = (!conv && !divg && a.prm.max_iter > 0) ? 1 : 0;
        sh.beta[b] = 0.f; sh.alpha[b] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int any = 0; for (int b = 0; b < batch; ++b) any |= sh.cont[b]; *sh.any_cont = any; }
    __syncthreads();

    float* dold = a.d0; float* dnew = a.d1;
    while (*sh.any_cont) {
        {   // pass A
            const float* hsrc[2] = {a.r, dold};
            const float* esrc[2] = {nullptr, nullptr};
            sweep(sh.cont, [&](const RingUnit& u, float& acc0, float& acc1) {
                REpiPassA epi{dnew, 0.f, 0.f};
                ring_process_unit<DIM, 2, 0>(rg, cfg, g, a.pf, hsrc, esrc, sh.beta[u.b], u, epi);
                acc0 += epi.acc0; acc1 += epi.acc1;
            });
        }
        barrier_and_reduce(sh.cont);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            if (!sh.cont[b]) continue;
            const double S = sh.sum1[b];
            const double dq = sh.sum0[b] + (double)coffs * S * S;
            sh.alpha[b] = (dq != 0.0) ? (float)(sh.delta[b] / dq) : 0.f;
            sh.offs[b] = coffs * (float)S;
        }
        __syncthreads();
        {   // pass B
            const float* hsrc[2] = {dnew, nullptr};
            const float* esrc[2] = {a.x, a.r};
            sweep(sh.cont, [&](const RingUnit& u, float& acc0, float& acc1) {
                REpiPassB epi{a.x, a.r, sh.alpha[u.b], sh.offs[u.b], 0.f, 0.f};
                ring_process_unit<DIM, 1, 2>(rg, cfg, g, a.pf, hsrc, esrc, 0.f, u, epi);
                acc0 += epi.acc0;
            });
        }
        barrier_and_reduce(sh.cont);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            if (!sh.cont[b]) continue;
            const double dn = sh.sum0[b];
            const double dprev = sh.delta[b];
            sh.beta[b] = (dprev != 0.0) ? (float)(dn / dprev) : 0.f;
            sh.delta[b] = dn;
            const int it = ++sh.iters[b];
            const float rsq = fabsf((float)dn);
            const bool conv = rsq <= sh.tol_sq[b];
            const bool divg = !isfinite(rsq) || (rsq / sh.rsq0[b] > 1e5f && it >= 8);
            sh.conv[b] = conv; sh.divg[b] = divg;
            sh.cont[b] = (!conv && !divg && it < a.prm.max_iter) ? 1 : 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int any = 0; for (int b = 0; b < batch; ++b) any |= sh.cont[b]; *sh.any_cont = any; }
        __syncthreads();
        float* t = dold; dold = dnew; dnew = t;
    }

    if (a.prm.project_mean) {
        sweep(nullptr, [&](const RingUnit& u, float& acc0, float& acc1) {
            ring_unit_cells<DIM>(cfg, g, a.pf, u, [&](long long off, int nvalid) {
                for (int j = 0; j < nvalid; ++j) acc0 += a.x[off + j];
            });
        });
        barrier_and_reduce(nullptr);
        for (int unit = blockIdx.x; unit < cfg.total_units; unit += gridDim.x) {
            const RingUnit u = ring_unit<DIM>(cfg, g, unit);
            const float m = (float)(sh.sum0[u.b] / cells);
            ring_unit_cells<DIM>(cfg, g, a.pf, u, [&](long long off, int nvalid) {
                for (int j = 0; j < nvalid; ++j) a.x[off + j] -= m;
            });
        }
    }

    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            PhiCgResult res;
            res.iterations = sh.iters[b]; res.converged = sh.conv[b]; res.diverged = sh.divg[b];
            res.residual_sq = fabsf((float)sh.delta[b]); res.tol_sq = sh.tol_sq[b]; res.initial_residual_sq = sh.rsq0[b];
            a.result[b] = res;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
static const int kSmemBudget = 227 * 1024;

// rows_per_ty(TY) = lines staged per stage; returns false when the grid lines are too long for a useful ring
static bool ring_config(const DGrid& g, int lines_a, int lines_b, int reserve_bytes, int min_stages, int max_stages,
                        int target_units, RingCfg* out)
{
    RingCfg c;
    c.pitch = g.cext[0]; c.nx4 = g.cext[0] / 4;
    const int row_bytes = c.pitch * 4;
    const int ty_max = g.dim == 3 ? 8 : 16;
    int ty = ty_max;
    for (;; ty /= 2) {
        if (ty < 1) return false;
        const int stage_bytes = (lines_a * ty + lines_b) * row_bytes;          // lines = lines_a*TY + lines_b
        const int r = (kSmemBudget - reserve_bytes - 128) / stage_bytes;
        if (r >= min_stages) { c.TY = ty; c.R = r > max_stages ? max_stages : r; c.stage_floats = stage_bytes / 4; break; }
    }
    while (c.TY > 1 && c.TY / 2 >= g.n[1] && g.dim == 2) c.TY /= 2;
    if (g.dim == 3) {
        c.nyt = (g.n[1] + c.TY - 1) / c.TY;
        int zc = g.n[2] < 64 ? g.n[2] : 64;
        for (;;) {
            c.ZC = zc; c.nzc = (g.n[2] + zc - 1) / zc;
            c.units_per_batch = c.nyt * c.nzc;
            c.total_units = c.units_per_batch * g.batch;
            if (c.total_units >= target_units || zc <= 8) break;
            zc /= 2;
        }
    } else {
        c.nyt = (g.n[1] + c.TY - 1) / c.TY; c.nzc = 1; c.ZC = 1;
        c.units_per_batch = c.nyt; c.total_units = c.nyt * g.batch;
    }
    *out = c;
    return true;
}

static int sm_count()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// returns -100 when the ring does not apply (caller falls back to the register-marching kernel)
int phi_launch_laplace_ring(const DGrid& g, const DField& f, const float* x, float* y, float coeff, bool axpy, cudaStream_t s)
{
    RingCfg cfg;
    const int sms = sm_count();
    // two CTAs per SM: each gets half of the shared memory
    if (!ring_config(g, 1, 2, kSmemBudget / 2 + 1024, g.dim == 3 ? 4 : 2, 6, sms * 2 * 4, &cfg)) return -100;
    const size_t smem = 128 + (size_t)cfg.R * cfg.stage_floats * 4;
    int grid = sms * 2;
    if (grid > cfg.total_units) grid = cfg.total_units;
    cudaError_t e;
#define LAUNCH_LAP(D, AX) do { \
        e = cudaFuncSetAttribute(k_laplace_ring<D, AX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) k_laplace_ring<D, AX><<<grid, RING_THREADS, smem, s>>>(g, f, cfg, x, y, coeff); } while (0)
    if (g.dim == 3) { if (axpy) LAUNCH_LAP(3, true); else LAUNCH_LAP(3, false); }
    else            { if (axpy) LAUNCH_LAP(2, true); else LAUNCH_LAP(2, false); }
#undef LAUNCH_LAP
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

int phi_launch_cg_ring(const CgLaunch& l, cudaStream_t s)
{
    const DGrid& g = l.g;
    if (g.batch > CG_MAX_BATCH) return -100;
    CgRingArgs A;
    const int cgs = (int)((cg_smem_bytes(g.batch) + 127) / 128 * 128);
    const int sms = sm_count();
    if (!ring_config(g, 3, 4, cgs, g.dim == 3 ? 4 : 2, RING_MAX_STAGES, sms * 6, &A.cfg)) return -100;
    A.ring_smem_offset = cgs;
    const size_t smem = (size_t)cgs + 128 + (size_t)A.cfg.R * A.cfg.stage_floats * 4;
    int per_sm = 0;
    cudaError_t e;
    if (g.dim == 3) {
        e = cudaFuncSetAttribute(k_cg_ring<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_ring<3>, RING_THREADS, smem);
    } else {
        e = cudaFuncSetAttribute(k_cg_ring<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_ring<2>, RING_THREADS, smem);
    }
    if (e != cudaSuccess || per_sm < 1) return -100;
    int grid = sms * per_sm;
    if (grid > A.cfg.total_units) grid = A.cfg.total_units;
    if (grid > CG_MAX_GRID) grid = CG_MAX_GRID;
    const size_t pf_sb = (size_t)g.cext[0] * g.cext[1] * g.cext[2];
    const size_t arr = ((size_t)pf_sb * g.batch * sizeof(float) + 255) / 256 * 256;
    unsigned char* ws = (unsigned char*)l.workspace;
    CgArgs& a = A.a;
    a.g = g; a.pf = l.pf; a.um = UnitMap();
    a.rhs = l.rhs; a.x = l.x;
    a.r = (float*)ws; a.d0 = (float*)(ws + arr); a.d1 = (float*)(ws + 2 * arr);
    a.partials = (double*)(ws + 3 * arr);
    a.result = l.result; a.prm = l.prm;
    void* args[] = {&A};
    if (g.dim == 3) e = cudaLaunchCooperativeKernel((void*)k_cg_ring<3>, dim3(grid), dim3(RING_THREADS), args, smem, s);
    else            e = cudaLaunchCooperativeKernel((void*)k_cg_ring<2>, dim3(grid), dim3(RING_THREADS), args, smem, s);
    if (e != cudaSuccess) { phi_set_error("cg ring: cooperative launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
