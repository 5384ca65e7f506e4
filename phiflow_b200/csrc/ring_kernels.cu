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
//
// The consumer loop is issue-bound if written naively (first ncu capture: 200+ instructions per 4 cells), so every
// thread precomputes the geometry of the <= 4 float4 groups it owns once per kernel, boundary flags once per tile, and
// the common case runs a branch-free path: 5 LDS.128 + 2 SHFL + ~40 FP32 ops + 1 STG.128 per group.
#include <cooperative_groups.h>
#include <type_traits>
#include "cg_common.cuh"
#include <cstdlib>
#include <cstring>
#include "launch.cuh"

namespace cg = cooperative_groups;

#define RING_CONSUMERS 256
#define RING_THREADS (RING_CONSUMERS + 32)
#define RING_MAX_STAGES 8
#define RING_G 4                 // float4 groups per consumer thread and plane (TY * nx4 <= RING_G * 256)

struct RingCfg {
    int TY, R, pitch, nx4;
    int stage_floats;          // floats between consecutive stages
    int ZC, nyt, nzc;
    int units_per_batch, total_units;
    int consumers;             // consumer threads per CTA (256 or 512); the producer warp follows them
    int groups;                // ceil(TY * nx4 / consumers)
    int shfl_ok;               // lanes of a warp own consecutive groups of one line -> x neighbours via shuffles
    int hint;                  // 1: element-wise lines are fetched with an L2 evict-first policy
    int merge;                 // 1: lines that are contiguous in memory travel in one bulk copy
    int dbg;                   // profiling only (PHICUDA_RING_DEBUG): 1 skip CG pass A, 2 skip pass B, 4 consumers skip the arithmetic
    // "tail split" decomposition of the CG kernel (3-D, batch 1, fewer y tiles than persistent CTAs): CTA c < nyt marches tile c
    // over planes [0, Zm); the remaining CTAs share the tails [Zm, nz) of all tiles, split_t tiles each.  Keeps every SM busy
    // when the tile count does not divide the CTA count (128 tiles on 148 SMs for 512-wide slabs).
    int split, Zm, split_t;
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
// same copy with an L2 evict-first policy: element-wise streams (x, r, rhs) are read once per pass, so they should not push the
// halo lines that neighbouring CTAs still need out of L2
__device__ __forceinline__ void bulk_g2s_stream(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- ring state (per thread, registers) ------------------------------------------------------------------------------
struct SlotIt {                 // a position in the ring: slot index + parity of its current use
    int slot; unsigned par;
    __device__ __forceinline__ void next(int R) { if (++slot == R) { slot = 0; par ^= 1u; } }
};

struct Ring {
    float* stage0;
    uint32_t stage0_s, full0, empty0;
    SlotIt pos;                // next plane to produce (producer) / first plane of the next unit (consumers)
};

__device__ __forceinline__ void ring_init(Ring& rg, unsigned char* smem, const RingCfg& cfg)
{
    // layout: [8 full barriers][8 empty barriers][stages]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    rg.full0 = smem_u32(bars);
    rg.empty0 = smem_u32(bars + RING_MAX_STAGES);
    rg.stage0 = reinterpret_cast<float*>(smem + 128);
    rg.stage0_s = smem_u32(rg.stage0);
    rg.pos.slot = 0; rg.pos.par = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cfg.R; ++s) {
            mbar_init(rg.full0 + 8 * s, 1);
            mbar_init(rg.empty0 + 8 * s, cfg.consumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// Producer warp.  Lane l owns staged lines l, l+32, l+64, l+96 of a stage: first the NH haloed arrays (TY+2 lines
// each), then the NE element-wise arrays (TY lines each).  Everything that does not depend on the plane is resolved once
// per unit (ProdUnit); per plane the producer only resolves z, adds it to the line offsets and issues the bulk copies -
// the producer is a single warp, so its serial instruction count per plane bounds the whole pipeline.
struct ProdUnit {
    long long yoff[4];          // b*sb + y*sy of the source line (boundary already applied to y)
    const float* base[4];       // source array
    uint32_t dsto[4];           // byte offset of the destination line inside a stage
    uint32_t nbytes[4];         // bytes of the copy this lane issues: consecutive lines of one array are merged into one copy
    unsigned hmask, emask;      // which of the 4 lines are active haloed / element-wise lines
    int tot_h, tot_e;           // warp totals of active lines
};

// hactive: bit k set = haloed slot k is fetched (slot 1 idles in the first CG iteration; with obstacles the last slot is the mask)
template <int DIM>
__device__ __forceinline__ void prod_unit_setup(ProdUnit& pu, const RingCfg& cfg, const DGrid& g, const DField& pf,
                                                unsigned hactive, int NHslots, int NE, const float* const* hsrc, const float* const* esrc,
                                                int b, int y0)
{
    const int lane = threadIdx.x & 31, hrows = cfg.TY + 2;
    const uint32_t row_bytes = (uint32_t)cfg.pitch * 4u;
    const bool mergeable = cfg.merge && pf.sy == cfg.pitch;      // neighbouring lines are contiguous in memory and in the stage
    pu.hmask = pu.emask = 0;
    int ch = 0, ce = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = lane + 32 * it;
        pu.yoff[it] = 0; pu.base[it] = nullptr; pu.dsto[it] = 0; pu.nbytes[it] = row_bytes;
        int key = -1, yv = 0;                                    // array of an active line, its source row
        bool halo_line = false;
        if (r < NHslots * hrows) {
            const int arr = r / hrows, j = r - arr * hrows;
            int yy = y0 - 1 + j; float cv;
            if (((hactive >> arr) & 1u) && yy <= g.n[1] && phi_resolve(yy, pf, 1, cv)) {
                pu.yoff[it] = (long long)b * pf.sb + (long long)yy * pf.sy;
                pu.base[it] = hsrc[arr];
                pu.dsto[it] = 4u * (uint32_t)((arr * hrows + j) * cfg.pitch);
                key = arr; yv = yy; halo_line = true; ch++;
            }
        } else if (r < NHslots * hrows + NE * cfg.TY) {
            const int q = r - NHslots * hrows, arr = q / cfg.TY, j = q - arr * cfg.TY;
            const int yy = y0 + j;
            if (yy < g.n[1]) {
                pu.yoff[it] = (long long)b * pf.sb + (long long)yy * pf.sy;
                pu.base[it] = esrc[arr];
                pu.dsto[it] = 4u * (uint32_t)((NHslots * hrows + arr * cfg.TY + j) * cfg.pitch);
                key = 8 + arr; yv = yy; ce++;
            }
        }
        // a line that continues the one held by the previous lane (same array, next row) rides on that lane's copy
        const int pkey = __shfl_up_sync(0xffffffffu, key, 1), pyv = __shfl_up_sync(0xffffffffu, yv, 1);
        const bool cont = mergeable && key >= 0 && lane > 0 && pkey == key && pyv + 1 == yv;
        const unsigned cm = __ballot_sync(0xffffffffu, cont);
        if (key >= 0 && !cont) {
            const unsigned follow = lane == 31 ? 0u : (cm >> (lane + 1));
            pu.nbytes[it] = (uint32_t)__ffs(~follow) * row_bytes;          // 1 + number of lines that continue this one
            if (halo_line) pu.hmask |= 1u << it; else pu.emask |= 1u << it;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ch += __shfl_xor_sync(0xffffffffu, ch, o); ce += __shfl_xor_sync(0xffffffffu, ce, o); }
    pu.tot_h = ch; pu.tot_e = ce;
}

template <int DIM>
__device__ __forceinline__ void ring_produce(Ring& rg, const RingCfg& cfg, const DField& pf, const ProdUnit& pu, int z, bool interior)
{
    const int lane = threadIdx.x & 31;
    const int slot = rg.pos.slot;
    const uint32_t full = rg.full0 + 8 * slot;
    const uint32_t row_bytes = (uint32_t)cfg.pitch * 4u;
    const uint32_t sbase = rg.stage0_s + 4u * (uint32_t)(slot * cfg.stage_floats);
    float cv;
    bool zmem = true;
    if (DIM == 3) zmem = phi_resolve(z, pf, 2, cv); else z = 0;
    const long long zoff = (long long)z * pf.sz;
    unsigned mask = (zmem ? pu.hmask : 0u) | (interior ? pu.emask : 0u);
    const int cnt = (zmem ? pu.tot_h : 0) + (interior ? pu.tot_e : 0);
    if (lane == 0) {
        mbar_wait(rg.empty0 + 8 * slot, rg.pos.par ^ 1u);
        if (cnt > 0) mbar_expect_tx(full, (uint32_t)cnt * row_bytes); else mbar_arrive(full);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it)
        if (mask & (1u << it)) {
            if (cfg.hint && (pu.emask & (1u << it))) bulk_g2s_stream(sbase + pu.dsto[it], pu.base[it] + pu.yoff[it] + zoff, pu.nbytes[it], full);
            else bulk_g2s(sbase + pu.dsto[it], pu.base[it] + pu.yoff[it] + zoff, pu.nbytes[it], full);
        }
    rg.pos.next(cfg.R);
}

__device__ __forceinline__ void ring_wait_full(const Ring& rg, const SlotIt& s) { mbar_wait(rg.full0 + 8 * s.slot, s.par); }
__device__ __forceinline__ void ring_release(const Ring& rg, const SlotIt& s)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(rg.empty0 + 8 * s.slot);
}
__device__ __forceinline__ const float* ring_ptr(const Ring& rg, const RingCfg& cfg, const SlotIt& s)
{ return rg.stage0 + (size_t)s.slot * cfg.stage_floats; }

// ---- consumer geometry --------------------------------------------------------------------------------------------------
#define GF_VALID 1u
#define GF_XLO   2u     // group holds x = 0
#define GF_XHI   4u     // group holds the last cell of the line
#define GF_YLOC  8u     // y-1 neighbour is a constant ghost
#define GF_YHIC  16u    // y+1 neighbour is a constant ghost

struct ThreadGroups {
    int soff[RING_G];       // float offset of the group inside a haloed array of a stage: (j+1)*pitch + x0
    int eoff[RING_G];       // float offset inside an element-wise array of a stage: j*pitch + x0
    int goff[RING_G];       // element offset relative to (y0, x=0) of the plane: j*sy + x0
    int j[RING_G];
    int xlo[RING_G], xro[RING_G];   // where the x-1 / x+4 neighbour lives in the staged line (ghosts resolved), fast path
    unsigned needl, needr;  // bit k: the neighbour of group k cannot come from a shuffle (warp edge or line end)
    unsigned flags[RING_G]; // per tile
    bool fast_ok;           // kernel-level eligibility of the branch-free path
};

__device__ __forceinline__ void groups_init(ThreadGroups& tg, const RingCfg& cfg, const DGrid& g, const DField& pf)
{
    const int lane = threadIdx.x & 31, nx = g.n[0];
    tg.needl = tg.needr = 0;
#pragma unroll
    for (int k = 0; k < RING_G; ++k) {
        const int gi = threadIdx.x + k * cfg.consumers;
        const int j = gi / cfg.nx4, x0 = (gi - j * cfg.nx4) * 4;
        tg.j[k] = (k < cfg.groups && j < cfg.TY && x0 < nx) ? j : -1;
        tg.soff[k] = (j + 1) * cfg.pitch + x0;
        tg.eoff[k] = j * cfg.pitch + x0;
        tg.goff[k] = j * (int)pf.sy + x0;
        tg.flags[k] = 0;
        const int row = (j + 1) * cfg.pitch;
        tg.xlo[k] = tg.soff[k] - 1; tg.xro[k] = tg.soff[k] + 4;
        if (lane == 0) tg.needl |= 1u << k;
        if (lane == 31) tg.needr |= 1u << k;
        if (x0 == 0) { tg.needl |= 1u << k; tg.xlo[k] = pf.klo[0] == PHI_BC_PERIODIC ? row + nx - 1 : row; }
        if (x0 + 4 >= nx) { tg.needr |= 1u << k; tg.xro[k] = pf.khi[0] == PHI_BC_PERIODIC ? row : row + nx - 1; }
        // lanes that take their neighbour from a shuffle still execute the (branch-free) edge load: one broadcast address
        if (!(tg.needl & (1u << k))) tg.xlo[k] = 0;
        if (!(tg.needr & (1u << k))) tg.xro[k] = 0;
    }
    tg.fast_ok = cfg.shfl_ok && (cfg.nx4 * 4 == nx) && (cfg.TY * cfg.nx4 == cfg.groups * cfg.consumers)
                 && pf.klo[0] != PHI_BC_CONST && pf.khi[0] != PHI_BC_CONST;
}

__device__ __forceinline__ void groups_tile(ThreadGroups& tg, const RingCfg& cfg, const DGrid& g, const DField& pf, int y0)
{
    const int nx = g.n[0], ny = g.n[1];
#pragma unroll
    for (int k = 0; k < RING_G; ++k) {
        unsigned f = 0;
        if (tg.j[k] >= 0) {
            const int y = y0 + tg.j[k];
            const int x0 = tg.eoff[k] - tg.j[k] * cfg.pitch;
            if (y < ny) {
                f = GF_VALID;
                if (x0 == 0) f |= GF_XLO;
                if (x0 + 4 >= nx) f |= GF_XHI;
                if (y == 0 && pf.klo[1] == PHI_BC_CONST) f |= GF_YLOC;
                if (y == ny - 1 && pf.khi[1] == PHI_BC_CONST) f |= GF_YHIC;
            }
        }
        tg.flags[k] = f;
    }
}

// ---- consumer: one plane of one tile ---------------------------------------------------------------------------------
// sm/sc/sp: stages holding planes z-1, z, z+1 (sm/sp unused in 2-D).  Values of the differenced array are h0 (+ beta*h1).
// MASK (N4, static obstacles): the last haloed slot of a stage holds the accessible mask; the stencil becomes
//   q_c = sum_faces min(acc_c, acc_nb) * (v_nb - v_c) / dx^2 for fluid cells, q_c = v_c inside obstacles
// (fluid.masked_laplace, phi/physics/fluid.py:197-202); constant (Dirichlet) ghost cells count as accessible.
template <int DIM, int NH, int NE, bool MASK = false, class Epi>
__device__ __forceinline__ void ring_compute(const RingCfg& cfg, const DGrid& g, const DField& pf, const ThreadGroups& tg,
                                             const float* sm, const float* sc, const float* sp, float beta,
                                             long long plane_off, int z, Epi& epi)
{
    const int pitch = cfg.pitch;
    const int h1 = (cfg.TY + 2) * pitch;                 // offset of the second haloed array inside a stage
    const int aoff = NH * (cfg.TY + 2) * pitch;          // MASK: offset of the accessible mask
    const int e0off = (NH + (MASK ? 1 : 0)) * (cfg.TY + 2) * pitch;
    const int e1off = e0off + cfg.TY * pitch;
    const float ix2 = g.inv_dx2[0], iy2 = g.inv_dx2[1], iz2 = g.inv_dx2[2];
    const bool zm_const = DIM == 3 && z == 0 && pf.klo[2] == PHI_BC_CONST;
    const bool zp_const = DIM == 3 && z == g.n[2] - 1 && pf.khi[2] == PHI_BC_CONST;
    const bool use1 = NH == 2 && beta != 0.f;
    const int lane = threadIdx.x & 31;

    auto val4 = [&](const float* s, int off) -> float4 {
        float4 a = *reinterpret_cast<const float4*>(s + off);
        if (use1) { const float4 o = *reinterpret_cast<const float4*>(s + h1 + off);
                    a.x = fmaf(beta, o.x, a.x); a.y = fmaf(beta, o.y, a.y); a.z = fmaf(beta, o.z, a.z); a.w = fmaf(beta, o.w, a.w); }
        return a;
    };
    auto val1 = [&](const float* s, int off) -> float {
        float a = s[off];
        if (use1) a = fmaf(beta, s[h1 + off], a);
        return a;
    };

#pragma unroll
    for (int k = 0; k < RING_G; ++k) {
        const unsigned f = tg.flags[k];
        const int rc = tg.soff[k];
        const bool valid = (f & GF_VALID) != 0;
        float4 c = f4_splat(0.f);
        if (valid) c = val4(sc, rc);
        float xl, xr;
        if (cfg.shfl_ok) {                                  // whole warp executes the shuffles (uniform branch)
            xl = __shfl_up_sync(0xffffffffu, c.w, 1);
            xr = __shfl_down_sync(0xffffffffu, c.x, 1);
            if (valid && lane == 0 && !(f & GF_XLO)) xl = val1(sc, rc - 1);
            if (valid && lane == 31 && !(f & GF_XHI)) xr = val1(sc, rc + 4);
        } else if (valid) {
            xl = (f & GF_XLO) ? 0.f : val1(sc, rc - 1);
            xr = (f & GF_XHI) ? 0.f : val1(sc, rc + 4);
        }
        if (!valid) continue;
        float4 ym, yp, q;
        int nvalid = 4;
        if (f == GF_VALID) {                                // interior of the line, no constant ghosts in y
            ym = val4(sc, rc - pitch);
            yp = val4(sc, rc + pitch);
        } else {
            const int row = rc - (tg.eoff[k] - tg.j[k] * pitch);          // start of the staged line
            const int nx = g.n[0];
            const int x0 = rc - row;
            nvalid = min(4, nx - x0);
            ym = (f & GF_YLOC) ? f4_splat(pf.clo[1]) : val4(sc, rc - pitch);
            yp = (f & GF_YHIC) ? f4_splat(pf.chi[1]) : val4(sc, rc + pitch);
            if (f & GF_XLO) { const int kx = pf.klo[0]; xl = kx == PHI_BC_PERIODIC ? val1(sc, row + nx - 1) : (kx == PHI_BC_ZERO_GRADIENT ? c.x : pf.clo[0]); }
            if (f & GF_XHI) { const int kx = pf.khi[0]; xr = kx == PHI_BC_PERIODIC ? val1(sc, row) : (kx == PHI_BC_ZERO_GRADIENT ? f4_get(c, nvalid - 1) : pf.chi[0]); }
        }
        float4 r4 = make_float4(c.y, c.z, c.w, xr);
        if (nvalid < 4) f4_set(r4, nvalid - 1, xr);
        if (MASK) {
            // the mask travels like the values: same staged lines, ghost cells 1 where the value ghost is a constant
            const float* am = sm + aoff; const float* ac_ = sc + aoff; const float* ap = sp + aoff;
            const float4 ac = *reinterpret_cast<const float4*>(ac_ + rc);
            const int row = rc - (tg.eoff[k] - tg.j[k] * pitch);
            const int nx = g.n[0];
            float axl = (f & GF_XLO) ? (pf.klo[0] == PHI_BC_PERIODIC ? ac_[row + nx - 1] : (pf.klo[0] == PHI_BC_ZERO_GRADIENT ? ac.x : 1.f)) : ac_[rc - 1];
            float axr = (f & GF_XHI) ? (pf.khi[0] == PHI_BC_PERIODIC ? ac_[row] : (pf.khi[0] == PHI_BC_ZERO_GRADIENT ? f4_get(ac, nvalid - 1) : 1.f)) : ac_[rc + 4];
            const float4 al4 = make_float4(axl, ac.x, ac.y, ac.z);
            float4 ar4 = make_float4(ac.y, ac.z, ac.w, axr);
            if (nvalid < 4) f4_set(ar4, nvalid - 1, axr);
            const float4 l4 = make_float4(xl, c.x, c.y, c.z);
            const float4 aym = (f & GF_YLOC) ? f4_splat(1.f) : *reinterpret_cast<const float4*>(ac_ + rc - pitch);
            const float4 ayp = (f & GF_YHIC) ? f4_splat(1.f) : *reinterpret_cast<const float4*>(ac_ + rc + pitch);
            auto term = [](float vn, float vc, float an, float a0) { return fminf(an, a0) * (vn - vc); };
            q.x = (term(l4.x, c.x, al4.x, ac.x) + term(r4.x, c.x, ar4.x, ac.x)) * ix2 + (term(ym.x, c.x, aym.x, ac.x) + term(yp.x, c.x, ayp.x, ac.x)) * iy2;
            q.y = (term(l4.y, c.y, al4.y, ac.y) + term(r4.y, c.y, ar4.y, ac.y)) * ix2 + (term(ym.y, c.y, aym.y, ac.y) + term(yp.y, c.y, ayp.y, ac.y)) * iy2;
            q.z = (term(l4.z, c.z, al4.z, ac.z) + term(r4.z, c.z, ar4.z, ac.z)) * ix2 + (term(ym.z, c.z, aym.z, ac.z) + term(yp.z, c.z, ayp.z, ac.z)) * iy2;
            q.w = (term(l4.w, c.w, al4.w, ac.w) + term(r4.w, c.w, ar4.w, ac.w)) * ix2 + (term(ym.w, c.w, aym.w, ac.w) + term(yp.w, c.w, ayp.w, ac.w)) * iy2;
            if (DIM == 3) {
                const float4 zm = zm_const ? f4_splat(pf.clo[2]) : val4(sm, rc);
                const float4 zp = zp_const ? f4_splat(pf.chi[2]) : val4(sp, rc);
                const float4 azm = zm_const ? f4_splat(1.f) : *reinterpret_cast<const float4*>(am + rc);
                const float4 azp = zp_const ? f4_splat(1.f) : *reinterpret_cast<const float4*>(ap + rc);
                q.x += (term(zm.x, c.x, azm.x, ac.x) + term(zp.x, c.x, azp.x, ac.x)) * iz2;
                q.y += (term(zm.y, c.y, azm.y, ac.y) + term(zp.y, c.y, azp.y, ac.y)) * iz2;
                q.z += (term(zm.z, c.z, azm.z, ac.z) + term(zp.z, c.z, azp.z, ac.z)) * iz2;
                q.w += (term(zm.w, c.w, azm.w, ac.w) + term(zp.w, c.w, azp.w, ac.w)) * iz2;
            }
            if (ac.x == 0.f) q.x = c.x;
            if (ac.y == 0.f) q.y = c.y;
            if (ac.z == 0.f) q.z = c.z;
            if (ac.w == 0.f) q.w = c.w;
            epi.set_acc(ac);
        } else {
        q.x = (xl + r4.x - 2.f * c.x) * ix2 + (ym.x + yp.x - 2.f * c.x) * iy2;
        q.y = (c.x + r4.y - 2.f * c.y) * ix2 + (ym.y + yp.y - 2.f * c.y) * iy2;
        q.z = (c.y + r4.z - 2.f * c.z) * ix2 + (ym.z + yp.z - 2.f * c.z) * iy2;
        q.w = (c.z + r4.w - 2.f * c.w) * ix2 + (ym.w + yp.w - 2.f * c.w) * iy2;
        if (DIM == 3) {
            const float4 zm = zm_const ? f4_splat(pf.clo[2]) : val4(sm, rc);
            const float4 zp = zp_const ? f4_splat(pf.chi[2]) : val4(sp, rc);
            q.x += (zm.x + zp.x - 2.f * c.x) * iz2;
            q.y += (zm.y + zp.y - 2.f * c.y) * iz2;
            q.z += (zm.z + zp.z - 2.f * c.z) * iz2;
            q.w += (zm.w + zp.w - 2.f * c.w) * iz2;
        }
        }
        float4 e0 = f4_splat(0.f), e1 = f4_splat(0.f), e2 = f4_splat(0.f);
        if (NE >= 1) e0 = *reinterpret_cast<const float4*>(sc + e0off + tg.eoff[k]);
        if (NE >= 2) e1 = *reinterpret_cast<const float4*>(sc + e1off + tg.eoff[k]);
        if (NE >= 3) e2 = *reinterpret_cast<const float4*>(sc + e1off + cfg.TY * pitch + tg.eoff[k]);
        epi(plane_off + tg.goff[k], c, q, nvalid, e0, e1, e2);
    }
}

// Branch-free variant for tiles that lie completely inside the grid in y, planes without constant z ghosts and
// non-constant x boundaries: every thread owns exactly G full groups.
// In 3-D a thread owns the same cells on every plane of a unit, so the (combined) values of planes z-1 and z are carried
// in registers from plane to plane (ZMarch) and only plane z+1 is read from shared memory.
struct ZMarch { float4 m[2], c[2]; bool have; };      // only used with <= 2 groups per thread (register budget)

template <int DIM, int NH, int NE, int G, bool MARCH, class Epi>
__device__ __forceinline__ void ring_compute_fast(const RingCfg& cfg, const DGrid& g, const ThreadGroups& tg,
                                                  const float* sm, const float* sc, const float* sp, float beta,
                                                  long long plane_off, Epi& epi, ZMarch& zs)
{
    const int pitch = cfg.pitch;
    const int h1 = (cfg.TY + 2) * pitch;
    const int e0off = NH * (cfg.TY + 2) * pitch;
    const int e1off = e0off + cfg.TY * pitch;
    const float ix2 = g.inv_dx2[0], iy2 = g.inv_dx2[1], iz2 = DIM == 3 ? g.inv_dx2[2] : 0.f;
    const float cc = 2.f * (ix2 + iy2 + iz2);
    const bool use1 = NH == 2 && beta != 0.f;
    auto val4 = [&](const float* s, int off) -> float4 {
        float4 a = *reinterpret_cast<const float4*>(s + off);
        if (use1) { const float4 o = *reinterpret_cast<const float4*>(s + h1 + off);
                    a.x = fmaf(beta, o.x, a.x); a.y = fmaf(beta, o.y, a.y); a.z = fmaf(beta, o.z, a.z); a.w = fmaf(beta, o.w, a.w); }
        return a;
    };
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const int rc = tg.soff[k];
        constexpr bool MZ = MARCH && DIM == 3 && G <= 2;
        const float4 c = (MZ && zs.have) ? zs.c[k & 1] : val4(sc, rc);
        const float4 ym = val4(sc, rc - pitch);
        const float4 yp = val4(sc, rc + pitch);
        float xl = __shfl_up_sync(0xffffffffu, c.w, 1);
        float xr = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (MARCH) {        // latency-bound single-CTA kernels: every lane loads (broadcast address for most), then selects
            float el = sc[tg.xlo[k]], er = sc[tg.xro[k]];
            if (use1) { el = fmaf(beta, sc[h1 + tg.xlo[k]], el); er = fmaf(beta, sc[h1 + tg.xro[k]], er); }
            xl = (tg.needl & (1u << k)) ? el : xl;
            xr = (tg.needr & (1u << k)) ? er : xr;
        } else {
            if (tg.needl & (1u << k)) { xl = sc[tg.xlo[k]]; if (use1) xl = fmaf(beta, sc[h1 + tg.xlo[k]], xl); }
            if (tg.needr & (1u << k)) { xr = sc[tg.xro[k]]; if (use1) xr = fmaf(beta, sc[h1 + tg.xro[k]], xr); }
        }
        float4 q;
        q.x = fmaf(ix2, xl + c.y, fmaf(iy2, ym.x + yp.x, -cc * c.x));
        q.y = fmaf(ix2, c.x + c.z, fmaf(iy2, ym.y + yp.y, -cc * c.y));
        q.z = fmaf(ix2, c.y + c.w, fmaf(iy2, ym.z + yp.z, -cc * c.z));
        q.w = fmaf(ix2, c.z + xr, fmaf(iy2, ym.w + yp.w, -cc * c.w));
        if (DIM == 3) {
            const float4 zm = (MZ && zs.have) ? zs.m[k & 1] : val4(sm, rc);
            const float4 zp = val4(sp, rc);
            q.x = fmaf(iz2, zm.x + zp.x, q.x);
            q.y = fmaf(iz2, zm.y + zp.y, q.y);
            q.z = fmaf(iz2, zm.z + zp.z, q.z);
            q.w = fmaf(iz2, zm.w + zp.w, q.w);
            if (MZ) { zs.m[k & 1] = c; zs.c[k & 1] = zp; }
        }
        float4 e0 = f4_splat(0.f), e1 = f4_splat(0.f), e2 = f4_splat(0.f);
        if (NE >= 1) e0 = *reinterpret_cast<const float4*>(sc + e0off + tg.eoff[k]);
        if (NE >= 2) e1 = *reinterpret_cast<const float4*>(sc + e1off + tg.eoff[k]);
        if (NE >= 3) e2 = *reinterpret_cast<const float4*>(sc + e1off + cfg.TY * pitch + tg.eoff[k]);
        epi(plane_off + tg.goff[k], c, q, 4, e0, e1, e2);
    }
    zs.have = MARCH && DIM == 3 && G <= 2;
}

template <bool GENERIC, int DIM, int NH, int NE, bool MARCH, bool MASK = false, class Epi>
__device__ __forceinline__ void ring_compute_any(const RingCfg& cfg, const DGrid& g, const DField& pf, const ThreadGroups& tg,
                                                 bool fast, const float* sm, const float* sc, const float* sp, float beta,
                                                 long long plane_off, int z, Epi& epi, ZMarch& zs)
{
    if (cfg.dbg & 4) return;
    if (MASK) { zs.have = false; ring_compute<DIM, NH, NE, true>(cfg, g, pf, tg, sm, sc, sp, beta, plane_off, z, epi); return; }
    if (!GENERIC || fast) {
        if (cfg.groups == 4) { ring_compute_fast<DIM, NH, NE, 4, MARCH>(cfg, g, tg, sm, sc, sp, beta, plane_off, epi, zs); return; }
        if (cfg.groups == 2) { ring_compute_fast<DIM, NH, NE, 2, MARCH>(cfg, g, tg, sm, sc, sp, beta, plane_off, epi, zs); return; }
        if (cfg.groups == 1) { ring_compute_fast<DIM, NH, NE, 1, MARCH>(cfg, g, tg, sm, sc, sp, beta, plane_off, epi, zs); return; }
    }
    zs.have = false;
    if (GENERIC) ring_compute<DIM, NH, NE, false>(cfg, g, pf, tg, sm, sc, sp, beta, plane_off, z, epi);
}

// ---- one unit: producer streams its planes, consumers march through them --------------------------------------------------
struct RingUnit { int b, y0, z0, z1; };

template <int DIM>
__device__ __forceinline__ RingUnit ring_unit(const RingCfg& cfg, const DGrid& g, int unit)
{
    RingUnit u;
    u.b = unit / cfg.units_per_batch;
    const int r = unit - u.b * cfg.units_per_batch;
    if (DIM == 3) { const int zc = r / cfg.nyt, yt = r - zc * cfg.nyt; u.y0 = yt * cfg.TY; u.z0 = zc * cfg.ZC; u.z1 = min(g.n[2], u.z0 + cfg.ZC); }
    else { u.y0 = r * cfg.TY; u.z0 = 0; u.z1 = 1; }
    return u;
}

// k-th unit of this CTA; false when it has no more
template <int DIM>
__device__ __forceinline__ bool ring_next_unit(const RingCfg& cfg, const DGrid& g, int k, RingUnit& u)
{
    if (!cfg.split) {
        const int unit = blockIdx.x + k * gridDim.x;
        if (unit >= cfg.total_units) return false;
        u = ring_unit<DIM>(cfg, g, unit);
        return true;
    }
    if ((int)blockIdx.x < cfg.nyt) {
        if (k > 0) return false;
        u.b = 0; u.y0 = blockIdx.x * cfg.TY; u.z0 = 0; u.z1 = cfg.Zm;
        return true;
    }
    const int t = ((int)blockIdx.x - cfg.nyt) * cfg.split_t + k;
    if (k >= cfg.split_t || t >= cfg.nyt) return false;
    u.b = 0; u.y0 = t * cfg.TY; u.z0 = cfg.Zm; u.z1 = g.n[2];
    return true;
}

// consumers, 3-D: march through the planes of one unit.  G > 0: every plane takes the branch-free path with G groups.
template <bool GENERIC, int NH, int NE, bool MARCH, int G, bool MASK = false, class Epi>
__device__ __forceinline__ void ring_consume_planes(Ring& rg, const RingCfg& cfg, const DGrid& g, const DField& pf, const ThreadGroups& tg,
                                                    bool tile_fast, float beta, long long plane_off, const RingUnit& u, Epi& epi)
{
    const int nz = u.z1 - u.z0;
    SlotIt a = rg.pos, bq = a; bq.next(cfg.R);
    SlotIt c2 = bq; c2.next(cfg.R);
    ZMarch zs; zs.have = false;
    ring_wait_full(rg, a); ring_wait_full(rg, bq);
    for (int zi = 0; zi < nz; ++zi) {
        const int z = u.z0 + zi;
        ring_wait_full(rg, c2);
        epi.set_plane(z, g.n[2]);
        if (G > 0) {
            if (!(cfg.dbg & 4))
                ring_compute_fast<3, NH, NE, (G > 0 ? G : 1), MARCH>(cfg, g, tg, ring_ptr(rg, cfg, a), ring_ptr(rg, cfg, bq), ring_ptr(rg, cfg, c2),
                                                                     beta, plane_off, epi, zs);
        } else {
            const bool fast = tile_fast && !(z == 0 && pf.klo[2] == PHI_BC_CONST) && !(z == g.n[2] - 1 && pf.khi[2] == PHI_BC_CONST);
            ring_compute_any<GENERIC, 3, NH, NE, MARCH, MASK>(cfg, g, pf, tg, fast, ring_ptr(rg, cfg, a), ring_ptr(rg, cfg, bq), ring_ptr(rg, cfg, c2),
                                                              beta, plane_off, z, epi, zs);
        }
        ring_release(rg, a);
        a = bq; bq = c2; c2.next(cfg.R);
        plane_off += pf.sz;
    }
    ring_release(rg, a); ring_release(rg, bq);
    rg.pos = c2;
}

// MASK: hsrc[NH] is the accessible mask; it occupies the haloed slot after the NH value arrays.
template <bool GENERIC, int DIM, int NH, int NE, bool MARCH = true, bool MASK = false, class Epi>
__device__ __forceinline__ void ring_process_unit(Ring& rg, const RingCfg& cfg, const DGrid& g, const DField& pf,
                                                  ThreadGroups& tg,
                                                  const float* const* hsrc, const float* const* esrc, float beta,
                                                  const RingUnit& u, Epi& epi)
{
    const bool producer = (int)threadIdx.x >= cfg.consumers;
    if (producer) {
        const int nh = (NH == 2 && beta == 0.f) ? 1 : NH;          // first CG iteration: d' = r, old direction not read
        ProdUnit pu;
        prod_unit_setup<DIM>(pu, cfg, g, pf, ((1u << nh) - 1u) | (MASK ? (1u << NH) : 0u), NH + (MASK ? 1 : 0), NE, hsrc, esrc, u.b, u.y0);
        if (DIM == 3) {
            const int nz = u.z1 - u.z0;
            for (int p = 0; p < nz + 2; ++p)
                ring_produce<DIM>(rg, cfg, pf, pu, u.z0 - 1 + p, p >= 1 && p <= nz);
        } else {
            ring_produce<DIM>(rg, cfg, pf, pu, 0, true);
        }
        return;
    }
    const int ny = g.n[1];
    const bool tile_fast = !MASK && (!GENERIC || tg.fast_ok && u.y0 + cfg.TY <= ny
                           && !(u.y0 == 0 && pf.klo[1] == PHI_BC_CONST) && !(u.y0 + cfg.TY == ny && pf.khi[1] == PHI_BC_CONST));
    if (GENERIC) groups_tile(tg, cfg, g, pf, u.y0);      // also needed on fast tiles: planes with constant z ghosts take the slow path
    long long plane_off = (long long)u.b * pf.sb + (long long)u.y0 * pf.sy + (DIM == 3 ? (long long)u.z0 * pf.sz : 0);
    if (DIM == 3) {
        // kernels that contain only the branch-free path pick the group count once per unit, not once per plane
        if (!GENERIC && cfg.groups == 2) ring_consume_planes<GENERIC, NH, NE, MARCH, 2>(rg, cfg, g, pf, tg, tile_fast, beta, plane_off, u, epi);
        else if (!GENERIC && cfg.groups == 4) ring_consume_planes<GENERIC, NH, NE, MARCH, 4>(rg, cfg, g, pf, tg, tile_fast, beta, plane_off, u, epi);
        else if (!GENERIC) ring_consume_planes<GENERIC, NH, NE, MARCH, 1>(rg, cfg, g, pf, tg, tile_fast, beta, plane_off, u, epi);
        else ring_consume_planes<GENERIC, NH, NE, MARCH, 0, MASK>(rg, cfg, g, pf, tg, tile_fast, beta, plane_off, u, epi);
    } else {
        const SlotIt a = rg.pos;
        ring_wait_full(rg, a);
        const float* sc = ring_ptr(rg, cfg, a);
        epi.set_plane(-1, 0);
        ZMarch zs; zs.have = false;
        ring_compute_any<GENERIC, DIM, NH, NE, MARCH, MASK>(cfg, g, pf, tg, tile_fast, sc, sc, sc, beta, plane_off, 0, epi, zs);
        ring_release(rg, a);
        rg.pos.next(cfg.R);
    }
}

// ---- epilogues (values of element-wise arrays arrive from shared memory) -------------------------------------------------
// Multi-GPU: an epilogue that updates a vector whose halo the neighbouring slab needs also stores the first / last owned
// plane into that neighbour's halo plane through a peer pointer (NVLink).  plo / phi are pre-offset so that the same
// element offset `off` addresses the right halo plane; zf says whether the current plane is the first (1) / last (2).
struct NoPeerHalo {            // single-GPU kernels: nothing to store, nothing to test
    int zf;
    __device__ __forceinline__ void set_plane(int, int) {}
    __device__ __forceinline__ void put4(long long, const float4&) const {}
    __device__ __forceinline__ void put1(long long, float) const {}
};

struct PeerHalo {
    float* plo; float* phi; int zf;
    __device__ __forceinline__ void set_plane(int z, int nz) { zf = ((z == 0 && plo) ? 1 : 0) | ((z == nz - 1 && phi) ? 2 : 0); }
    __device__ __forceinline__ void put4(long long off, const float4& v) const
    {
        if (zf & 1) *reinterpret_cast<float4*>(plo + off) = v;
        if (zf & 2) *reinterpret_cast<float4*>(phi + off) = v;
    }
    __device__ __forceinline__ void put1(long long off, float v) const
    {
        if (zf & 1) plo[off] = v;
        if (zf & 2) phi[off] = v;
    }
};

template <bool AXPY>
struct REpiLaplace {
    float* y; float coeff;
    __device__ __forceinline__ void set_plane(int, int) {}
    __device__ __forceinline__ void set_acc(const float4&) {}
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4&, const float4&, const float4&)
    {
        float4 o = q;
        if (AXPY) { o.x = c.x + coeff * q.x; o.y = c.y + coeff * q.y; o.z = c.z + coeff * q.z; o.w = c.w + coeff * q.w; }
        if (nvalid == 4) *reinterpret_cast<float4*>(y + off) = o;
        else for (int j = 0; j < nvalid; ++j) y[off + j] = f4_get(o, j);
    }
};

template <class PH>
struct REpiResidual0 {          // e0 = rhs
    float* r; float mean, offs; float acc0, acc1; PH ph; bool tol_from_y;   // CG-adaptive: tolerance relative to |y|^2
    float4 am = make_float4(1.f, 1.f, 1.f, 1.f);      // obstacles: balanced rhs = y - mean * accessible (fluid.py:205-209)
    __device__ __forceinline__ void set_plane(int z, int nz) { ph.set_plane(z, nz); }
    __device__ __forceinline__ void set_acc(const float4& a) { am = a; }
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4& y, const float4&, const float4&)
    {
        float4 rt = make_float4((y.x - mean * am.x) - q.x, (y.y - mean * am.y) - q.y, (y.z - mean * am.z) - q.z, (y.w - mean * am.w) - q.w);
        float4 rr = make_float4(rt.x - offs, rt.y - offs, rt.z - offs, rt.w - offs);
        if (tol_from_y) rt = make_float4(y.x - mean * am.x, y.y - mean * am.y, y.z - mean * am.z, y.w - mean * am.w);
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(r + off) = rr;
            if (ph.zf) ph.put4(off, rr);
            acc0 += rr.x * rr.x + rr.y * rr.y + rr.z * rr.z + rr.w * rr.w;
            acc1 += rt.x * rt.x + rt.y * rt.y + rt.z * rt.z + rt.w * rt.w;
        } else for (int j = 0; j < nvalid; ++j) { const float a = f4_get(rr, j), t = f4_get(rt, j); r[off + j] = a; if (ph.zf) ph.put1(off + j, a); acc0 += a * a; acc1 += t * t; }
    }
};

template <class PH>
struct REpiPassA {
    __device__ __forceinline__ void set_acc(const float4&) {}
    float* dnew; float acc0, acc1; PH ph;
    __device__ __forceinline__ void set_plane(int z, int nz) { ph.set_plane(z, nz); }
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4&, const float4&, const float4&)
    {
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(dnew + off) = c;
            if (ph.zf) ph.put4(off, c);
            acc0 += c.x * q.x + c.y * q.y + c.z * q.z + c.w * q.w;
            acc1 += (c.x + c.y) + (c.z + c.w);
        } else for (int j = 0; j < nvalid; ++j) { const float v = f4_get(c, j); dnew[off + j] = v; if (ph.zf) ph.put1(off + j, v); acc0 += v * f4_get(q, j); acc1 += v; }
    }
};

// CG-adaptive pass A: the second sum is d'.r (e0 = r, element-wise), _linalg.py:113
template <class PH>
struct REpiPassAAdapt {
    __device__ __forceinline__ void set_acc(const float4&) {}
    float* dnew; float acc0, acc1; PH ph;
    __device__ __forceinline__ void set_plane(int z, int nz) { ph.set_plane(z, nz); }
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4& re, const float4&, const float4&)
    {
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(dnew + off) = c;
            if (ph.zf) ph.put4(off, c);
            acc0 += c.x * q.x + c.y * q.y + c.z * q.z + c.w * q.w;
            acc1 += c.x * re.x + c.y * re.y + c.z * re.z + c.w * re.w;
        } else for (int j = 0; j < nvalid; ++j) { const float v = f4_get(c, j); dnew[off + j] = v; if (ph.zf) ph.put1(off + j, v); acc0 += v * f4_get(q, j); acc1 += v * f4_get(re, j); }
    }
};

// The solution update is applied every second iteration only: x_{k+1} = x_{k-1} + alpha_{k-1} d_{k-1} + alpha_k d_k needs
// the previous direction (still intact in the other d buffer) but saves one read+write of x: 30 instead of 32 B/cell/it.
template <class PH, bool RQ = false>      // RQ (CG-adaptive): the second sum is r_new . q (_linalg.py:119)
struct REpiPassBr {
    __device__ __forceinline__ void set_acc(const float4&) {}             // odd iterations: e0 = r; x is left alone
    float* r; float alpha, offs; float acc0, acc1; PH ph;
    __device__ __forceinline__ void set_plane(int z, int nz) { ph.set_plane(z, nz); }
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4& re, const float4&, const float4&)
    {
        float4 rv = re;
        rv.x -= alpha * (q.x + offs); rv.y -= alpha * (q.y + offs); rv.z -= alpha * (q.z + offs); rv.w -= alpha * (q.w + offs);
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(r + off) = rv;
            if (ph.zf) ph.put4(off, rv);
            acc0 += rv.x * rv.x + rv.y * rv.y + rv.z * rv.z + rv.w * rv.w;
            if (RQ) acc1 += rv.x * q.x + rv.y * q.y + rv.z * q.z + rv.w * q.w;
        } else for (int j = 0; j < nvalid; ++j) { const float t = f4_get(rv, j); r[off + j] = t; if (ph.zf) ph.put1(off + j, t); acc0 += t * t; if (RQ) acc1 += t * f4_get(q, j); }
    }
};

template <class PH, bool RQ = false>
struct REpiPassB {
    __device__ __forceinline__ void set_acc(const float4&) {}              // even iterations: e0 = x, e1 = r, e2 = previous direction
    float* x; float* r; float alpha, aprev, offs; float acc0, acc1; PH ph;
    __device__ __forceinline__ void set_plane(int z, int nz) { ph.set_plane(z, nz); }
    __device__ __forceinline__ void operator()(long long off, const float4& c, const float4& q, int nvalid, const float4& xe, const float4& re, const float4& dp)
    {
        float4 xv = xe, rv = re;
        xv.x += aprev * dp.x; xv.y += aprev * dp.y; xv.z += aprev * dp.z; xv.w += aprev * dp.w;
        xv.x += alpha * c.x; xv.y += alpha * c.y; xv.z += alpha * c.z; xv.w += alpha * c.w;
        rv.x -= alpha * (q.x + offs); rv.y -= alpha * (q.y + offs); rv.z -= alpha * (q.z + offs); rv.w -= alpha * (q.w + offs);
        if (nvalid == 4) {
            *reinterpret_cast<float4*>(x + off) = xv;
            *reinterpret_cast<float4*>(r + off) = rv;
            if (ph.zf) ph.put4(off, rv);
            acc0 += rv.x * rv.x + rv.y * rv.y + rv.z * rv.z + rv.w * rv.w;
            if (RQ) acc1 += rv.x * q.x + rv.y * q.y + rv.z * q.z + rv.w * q.w;
        } else for (int j = 0; j < nvalid; ++j) { x[off + j] = f4_get(xv, j); const float t = f4_get(rv, j); r[off + j] = t; if (ph.zf) ph.put1(off + j, t); acc0 += t * t; if (RQ) acc1 += t * f4_get(q, j); }
    }
};

// ---- laplace -----------------------------------------------------------------------------------------------------------
template <int DIM, bool AXPY, bool GENERIC>
__global__ void __launch_bounds__(RING_THREADS, 2)
k_laplace_ring(DGrid g, DField f, RingCfg cfg, const float* __restrict__ x, float* __restrict__ y, float coeff)
{
    extern __shared__ __align__(128) unsigned char smem[];
    Ring rg;
    ring_init(rg, smem, cfg);
    ThreadGroups tg;
    groups_init(tg, cfg, g, f);
    const float* hsrc[2] = {x, nullptr};
    const float* esrc[2] = {nullptr, nullptr};
    REpiLaplace<AXPY> epi{y, coeff};
    for (int unit = blockIdx.x; unit < cfg.total_units; unit += gridDim.x) {
        const RingUnit u = ring_unit<DIM>(cfg, g, unit);
        ring_process_unit<GENERIC, DIM, 1, 0, false>(rg, cfg, g, f, tg, hsrc, esrc, 0.f, u, epi);
    }
}

// ---- CG ----------------------------------------------------------------------------------------------------------------
struct CgRingArgs {
    CgArgs a;
    RingCfg cfg;
    int ring_smem_offset;        // byte offset of the ring inside dynamic shared memory (after the CgShared block)
    int comm_merge;              // multi-GPU: 1 = merged barrier + all-reduce (comm_barrier_allreduce), 0 = grid.sync + comm_allreduce
    CommDev cm;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// All-reduce of the two per-batch sums across the ranks, executed inside the persistent kernel after the local
// reduction: block 0 writes this rank's sums into every rank's mailbox (NVLink peer stores) and raises a flag; every CTA
// of every rank waits for the n flags in its OWN mailbox and adds the n entries in rank order (bitwise identical
// everywhere).  Event `seq` uses mailbox half seq&1; a rank can be at most one event ahead of the slowest one.
__device__ __forceinline__ bool comm_allreduce(const CommDev& cm, const CgShared& sh, int batch, unsigned long long seq)
{
    const int par = (int)(seq & 1ull);
    const size_t stride = 2 * (size_t)CG_MAX_BATCH;
    if (blockIdx.x == 0) {
        for (int q = 0; q < cm.n; ++q) {
            double* dst = cm.mbox[q] + ((size_t)par * PHI_MAX_RANKS + cm.rank) * stride;
            for (int b = threadIdx.x; b < batch; b += blockDim.x) { dst[b] = sh.sum0[b]; dst[CG_MAX_BATCH + b] = sh.sum1[b]; }
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < cm.n) st_release_sys(cm.flag[threadIdx.x] + par * PHI_MAX_RANKS + cm.rank, seq);
    }
    bool ok = true;
    if ((int)threadIdx.x < cm.n) {
        const unsigned long long* f = cm.flag[cm.rank] + par * PHI_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > 40000000000ll) { ok = false; break; }        // ~20 s: a peer died; do not hang the GPU
        }
    }
    ok = __syncthreads_and(ok);
    const double* src = cm.mbox[cm.rank] + (size_t)par * PHI_MAX_RANKS * stride;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        double s0 = 0, s1 = 0;
        for (int q = 0; q < cm.n; ++q) {
            s0 += *(volatile const double*)(src + q * stride + b);
            s1 += *(volatile const double*)(src + q * stride + CG_MAX_BATCH + b);
        }
        sh.sum0[b] = s0; sh.sum1[b] = s1;
    }
    __syncthreads();
    return ok;
}

// Merged barrier + all-reduce of the multi-GPU kernels - an experiment, opt-in with PHICUDA_COMM_MERGE=1 (default: grid.sync +
// comm_allreduce; the two measure the same, see phi_launch_cg_ring).
// Every CTA makes its stores visible system-wide (partial sums, halo planes stored into the neighbours' memory) and arrives on a
// local counter; the LAST CTA to arrive sums the per-CTA partials in a fixed order and sends the result with a release flag to
// every rank (itself included); every CTA of every rank then waits for the n flags in its own mailbox and adds the n entries in
// rank order.  One arrival + one flag hop instead of a full grid barrier followed by block 0's send: the local release of the
// grid barrier, the redundant partial reduction in all CTAs and block 0's serial send are gone.  The own rank's flag is only
// written after all local CTAs arrived, so passing the wait still implies "every local CTA finished the pass".
__device__ __forceinline__ bool comm_barrier_allreduce(const CommDev& cm, const CgShared& sh, const double* partials, int region, int batch,
                                                       int units_per_batch, const unsigned char* active, unsigned long long seq)
{
    const int par = (int)(seq & 1ull);
    const size_t stride = 2 * (size_t)CG_MAX_BATCH;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(cm.arrive, 1u);
        sh.any_cont[1] = ((prev + 1u) % gridDim.x == 0u) ? 1 : 0;
    }
    __syncthreads();
    if (sh.any_cont[1]) {                                    // CTA-uniform: the last arriver
        __threadfence();
        reduce_partials(sh, partials, region, batch, units_per_batch, active);
        for (int q = 0; q < cm.n; ++q) {
            double* dst = cm.mbox[q] + ((size_t)par * PHI_MAX_RANKS + cm.rank) * stride;
            for (int b = threadIdx.x; b < batch; b += blockDim.x) { dst[b] = sh.sum0[b]; dst[CG_MAX_BATCH + b] = sh.sum1[b]; }
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < cm.n) st_release_sys(cm.flag[threadIdx.x] + par * PHI_MAX_RANKS + cm.rank, seq);
    }
    bool ok = true;
    if ((int)threadIdx.x < cm.n) {
        const unsigned long long* f = cm.flag[cm.rank] + par * PHI_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > 40000000000ll) { ok = false; break; }        // ~20 s: a peer died; do not hang the GPU
        }
    }
    ok = __syncthreads_and(ok);
    const double* src = cm.mbox[cm.rank] + (size_t)par * PHI_MAX_RANKS * stride;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        double s0 = 0, s1 = 0;
        for (int q = 0; q < cm.n; ++q) {
            s0 += *(volatile const double*)(src + q * stride + b);
            s1 += *(volatile const double*)(src + q * stride + CG_MAX_BATCH + b);
        }
        sh.sum0[b] = s0; sh.sum1[b] = s1;
    }
    __syncthreads();
    return ok;
}

template <int DIM, class F>
__device__ __forceinline__ void ring_unit_cells(const RingCfg& cfg, const DGrid& g, const DField& pf, const ThreadGroups& tg,
                                                const RingUnit& u, F&& fn)
{
    // plain element-wise traversal of a unit by the consumer threads (sums, mean removal); no staging
    if ((int)threadIdx.x >= cfg.consumers) return;
    long long plane_off = (long long)u.b * pf.sb + (long long)u.y0 * pf.sy + (DIM == 3 ? (long long)u.z0 * pf.sz : 0);
    for (int z = u.z0; z < u.z1; ++z, plane_off += pf.sz)
#pragma unroll
        for (int k = 0; k < RING_G; ++k) {
            if (tg.j[k] < 0 || u.y0 + tg.j[k] >= g.n[1]) continue;
            const int x0 = tg.eoff[k] - tg.j[k] * cfg.pitch;
            fn(plane_off + tg.goff[k], min(4, g.n[0] - x0));
        }
}

// MASK (N4): static obstacles - a.acc is the accessible mask, staged as an extra haloed array (always the GENERIC consumer).
template <int DIM, bool GENERIC, bool DIST, bool ADAPT, bool MASK = false>
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
    const double cells = (double)g.n[0] * g.n[1] * g.n[2] * (A.cm.n > 1 ? A.cm.n : 1);   // global cell count (equal slabs)
    const float coffs = a.prm.matrix_offset;
    constexpr bool adaptive = ADAPT;                 // a.prm.method == PHI_SOLVER_CG_ADAPTIVE, chosen by the launcher
    int region = 0;
    ThreadGroups tg;
    groups_init(tg, cfg, g, a.pf);

    auto sweep = [&](const unsigned char* active, auto&& body) {
        int cur_b = cfg.split ? 0 : -1; float acc0 = 0.f, acc1 = 0.f;      // split mode: every CTA reports for the one batch entry
        RingUnit u;
        for (int k = 0; ring_next_unit<DIM>(cfg, g, k, u); ++k) {
            if (active && !active[u.b]) continue;
            if (u.b != cur_b) {
                if (cur_b >= 0) flush_partials(sh, a.partials, region, batch, cur_b, acc0, acc1);
                cur_b = u.b; acc0 = 0.f; acc1 = 0.f;
            }
            body(u, acc0, acc1);
        }
        if (cur_b >= 0) flush_partials(sh, a.partials, region, batch, cur_b, acc0, acc1);
    };
    const CommDev& cm = A.cm;
    unsigned long long seq = cm.n > 1 ? *cm.seq : 0ull;
    bool comm_ok = true;
    auto barrier_and_reduce = [&](const unsigned char* active) {
        fence_proxy_async();                       // generic-proxy stores of this pass -> later TMA (async proxy) loads
        if (DIST && cm.n > 1 && A.comm_merge) {    // merged barrier + all-reduce
            comm_ok = comm_barrier_allreduce(cm, sh, a.partials, region, batch, cfg.split ? (int)gridDim.x : cfg.units_per_batch, active, ++seq) && comm_ok;
            region ^= 1;
            fence_proxy_async();
            return;
        }
        if (cm.n > 1) __threadfence_system();      // halo planes stored into the neighbours' memory
        grid.sync();
        fence_proxy_async();
        reduce_partials(sh, a.partials, region, batch, cfg.split ? (int)gridDim.x : cfg.units_per_batch, active);
        region ^= 1;
        if (cm.n > 1) {
            comm_ok = comm_allreduce(cm, sh, batch, ++seq) && comm_ok;
            fence_proxy_async();
        }
    };
    const long long nzsz = (long long)g.n[2] * a.pf.sz;
    using PH = typename std::conditional<DIST, PeerHalo, NoPeerHalo>::type;
    auto peer_halo = [&](float* lo_arr, float* hi_arr) {
        PH ph;
        if constexpr (DIST) {
            ph.plo = cm.lower >= 0 ? lo_arr + nzsz : nullptr;
            ph.phi = cm.upper >= 0 ? hi_arr - nzsz : nullptr;
        }
        ph.zf = 0;
        return ph;
    };

    for (int b = threadIdx.x; b < batch; b += blockDim.x) { sh.mean[b] = 0.f; sh.offs[b] = 0.f; }
    __syncthreads();

    if (a.prm.balance_rhs || coffs != 0.f) {
        sweep(nullptr, [&](const RingUnit& u, float& acc0, float& acc1) {
            ring_unit_cells<DIM>(cfg, g, a.pf, tg, u, [&](long long off, int nvalid) {
                for (int j = 0; j < nvalid; ++j) { acc0 += a.rhs[off + j]; acc1 += MASK ? a.acc[off + j] : a.x[off + j]; }
            });
        });
        barrier_and_reduce(nullptr);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            if (MASK) { sh.mean[b] = (a.prm.balance_rhs && sh.sum1[b] > 0.0) ? (float)(sh.sum0[b] / sh.sum1[b]) : 0.f; sh.offs[b] = 0.f; }
            else { sh.mean[b] = a.prm.balance_rhs ? (float)(sh.sum0[b] / cells) : 0.f; sh.offs[b] = coffs * (float)sh.sum1[b]; }
        }
        __syncthreads();
    }

    {   // r0 = y - (A + c 11^T) x0
        const float* hsrc[2] = {a.x, MASK ? a.acc : nullptr};
        const float* esrc[2] = {a.rhs, nullptr};
        sweep(nullptr, [&](const RingUnit& u, float& acc0, float& acc1) {
            REpiResidual0<PH> epi{a.r, sh.mean[u.b], sh.offs[u.b], 0.f, 0.f, peer_halo(cm.lo_r, cm.hi_r), adaptive};
            ring_process_unit<GENERIC, DIM, 1, 1, true, MASK>(rg, cfg, g, a.pf, tg, hsrc, esrc, 0.f, u, epi);
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
        sh.beta[b] = 0.f; sh.alpha[b] = 0.f; sh.aprev[b] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int any = 0; for (int b = 0; b < batch; ++b) any |= sh.cont[b]; *sh.any_cont = any; }
    __syncthreads();

    // a.prm.method == PHI_SOLVER_CG_ADAPTIVE (_linalg.py:93-128) reuses both passes: pass A forms d' = r - c d (beta = -c) and
    // sums d'.Ad' and d'.r, pass B applies the step (d'.r)/(d'.Ad') and sums |r|^2 and r.Ad' for the next c.
    float* dold = a.d0; float* dnew = a.d1;
    float* lo_dnew = cm.lo_d1; float* hi_dnew = cm.hi_d1; float* lo_dold = cm.lo_d0; float* hi_dold = cm.hi_d0;
    bool x_pending = false;      // all running entries are at the same iteration, so one flag describes them all
    while (*sh.any_cont && comm_ok) {
        if (cfg.dbg & 1) {} else if constexpr (!ADAPT) {   // pass A
            const float* hsrc[3] = {a.r, dold, MASK ? a.acc : nullptr};
            const float* esrc[2] = {nullptr, nullptr};
            sweep(sh.cont, [&](const RingUnit& u, float& acc0, float& acc1) {
                REpiPassA<PH> epi{dnew, 0.f, 0.f, peer_halo(lo_dnew, hi_dnew)};
                ring_process_unit<GENERIC, DIM, 2, 0, true, MASK>(rg, cfg, g, a.pf, tg, hsrc, esrc, sh.beta[u.b], u, epi);
                acc0 += epi.acc0; acc1 += epi.acc1;
            });
        } else {                                    // pass A of CG-adaptive: additionally d'.r (r once more, element-wise)
            const float* hsrc[2] = {a.r, dold};
            const float* esrc[2] = {a.r, nullptr};
            sweep(sh.cont, [&](const RingUnit& u, float& acc0, float& acc1) {
                REpiPassAAdapt<PH> epi{dnew, 0.f, 0.f, peer_halo(lo_dnew, hi_dnew)};
                ring_process_unit<GENERIC, DIM, 2, 1>(rg, cfg, g, a.pf, tg, hsrc, esrc, sh.beta[u.b], u, epi);
                acc0 += epi.acc0; acc1 += epi.acc1;
            });
        }
        barrier_and_reduce(sh.cont);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            if (!sh.cont[b]) continue;
            sh.aprev[b] = sh.alpha[b];
            if (adaptive) {                          // step = (d.r) / (d.Ad), divide_no_nan (_linalg.py:113-114)
                const double dq = sh.sum0[b];
                sh.alpha[b] = (dq != 0.0) ? (float)(sh.sum1[b] / dq) : 0.f;
                sh.delta[b] = dq;                    // kept for the direction update after pass B
                sh.offs[b] = 0.f;
                continue;
            }
            const double S = sh.sum1[b];
            const double dq = sh.sum0[b] + (double)coffs * S * S;
            sh.alpha[b] = (dq != 0.0) ? (float)(sh.delta[b] / dq) : 0.f;
            sh.offs[b] = coffs * (float)S;
        }
        __syncthreads();
        if (cfg.dbg & 2) {} else if (!x_pending) {   // pass B, odd iteration: r only, the x update is deferred
            const float* hsrc[2] = {dnew, MASK ? a.acc : nullptr};
            const float* esrc[3] = {a.r, nullptr, nullptr};
            sweep(sh.cont, [&](const RingUnit& u, float& acc0, float& acc1) {
                if constexpr (ADAPT) {
                    REpiPassBr<PH, true> epi{a.r, sh.alpha[u.b], 0.f, 0.f, 0.f, peer_halo(cm.lo_r, cm.hi_r)};
                    ring_process_unit<GENERIC, DIM, 1, 1>(rg, cfg, g, a.pf, tg, hsrc, esrc, 0.f, u, epi);
                    acc0 += epi.acc0; acc1 += epi.acc1;
                } else {
                    REpiPassBr<PH> epi{a.r, sh.alpha[u.b], sh.offs[u.b], 0.f, 0.f, peer_halo(cm.lo_r, cm.hi_r)};
                    ring_process_unit<GENERIC, DIM, 1, 1, true, MASK>(rg, cfg, g, a.pf, tg, hsrc, esrc, 0.f, u, epi);
                    acc0 += epi.acc0;
                }
            });
        } else {            // pass B, even iteration: x += alpha_prev d_prev + alpha d
            const float* hsrc[2] = {dnew, MASK ? a.acc : nullptr};
            const float* esrc[3] = {a.x, a.r, dold};
            sweep(sh.cont, [&](const RingUnit& u, float& acc0, float& acc1) {
                if constexpr (ADAPT) {
                    REpiPassB<PH, true> epi{a.x, a.r, sh.alpha[u.b], sh.aprev[u.b], 0.f, 0.f, 0.f, peer_halo(cm.lo_r, cm.hi_r)};
                    ring_process_unit<GENERIC, DIM, 1, 3>(rg, cfg, g, a.pf, tg, hsrc, esrc, 0.f, u, epi);
                    acc0 += epi.acc0; acc1 += epi.acc1;
                } else {
                    REpiPassB<PH> epi{a.x, a.r, sh.alpha[u.b], sh.aprev[u.b], sh.offs[u.b], 0.f, 0.f, peer_halo(cm.lo_r, cm.hi_r)};
                    ring_process_unit<GENERIC, DIM, 1, 3, true, MASK>(rg, cfg, g, a.pf, tg, hsrc, esrc, 0.f, u, epi);
                    acc0 += epi.acc0;
                }
            });
        }
        x_pending = !x_pending;
        barrier_and_reduce(sh.cont);
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            if (!sh.cont[b]) continue;
            const double dn = sh.sum0[b];
            const double dprev = sh.delta[b];                 // CG: previous |r|^2;  CG-adaptive: d.Ad of this iteration
            // CG: d' = r + (|r'|^2 / |r|^2) d;  CG-adaptive: d' = r - ((r.Ad) / (d.Ad)) d  (_linalg.py:120)
            if (adaptive) sh.beta[b] = (dprev != 0.0) ? -(float)(sh.sum1[b] / dprev) : 0.f;
            else sh.beta[b] = (dprev != 0.0) ? (float)(dn / dprev) : 0.f;
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
        t = lo_dold; lo_dold = lo_dnew; lo_dnew = t;
        t = hi_dold; hi_dold = hi_dnew; hi_dnew = t;
    }

    // entries that stopped after an odd number of iterations still owe x their last step; odd iterations write d1
    RingUnit u;
    for (int k = 0; ring_next_unit<DIM>(cfg, g, k, u); ++k) {
        if (!(sh.iters[u.b] & 1)) continue;
        const float al = sh.alpha[u.b];
        ring_unit_cells<DIM>(cfg, g, a.pf, tg, u, [&](long long off, int nvalid) {
            if (nvalid == 4) {
                float4 xv = *reinterpret_cast<const float4*>(a.x + off);
                const float4 dv = *reinterpret_cast<const float4*>(a.d1 + off);
                xv.x += al * dv.x; xv.y += al * dv.y; xv.z += al * dv.z; xv.w += al * dv.w;
                *reinterpret_cast<float4*>(a.x + off) = xv;
            } else for (int j = 0; j < nvalid; ++j) a.x[off + j] += al * a.d1[off + j];
        });
    }

    if (a.prm.project_mean) {
        sweep(nullptr, [&](const RingUnit& u, float& acc0, float& acc1) {
            ring_unit_cells<DIM>(cfg, g, a.pf, tg, u, [&](long long off, int nvalid) {
                for (int j = 0; j < nvalid; ++j) { if (MASK) { acc0 += a.x[off + j] * a.acc[off + j]; acc1 += a.acc[off + j]; } else acc0 += a.x[off + j]; }
            });
        });
        barrier_and_reduce(nullptr);
        for (int k = 0; ring_next_unit<DIM>(cfg, g, k, u); ++k) {
            const float m = MASK ? (sh.sum1[u.b] > 0.0 ? (float)(sh.sum0[u.b] / sh.sum1[u.b]) : 0.f) : (float)(sh.sum0[u.b] / cells);
            ring_unit_cells<DIM>(cfg, g, a.pf, tg, u, [&](long long off, int nvalid) {
                for (int j = 0; j < nvalid; ++j) a.x[off + j] -= MASK ? m * a.acc[off + j] : m;
            });
        }
    }

    if (cm.n > 1 && blockIdx.x == 0 && threadIdx.x == 0) *cm.seq = seq;
    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            PhiCgResult res;
            res.iterations = sh.iters[b]; res.converged = sh.conv[b]; res.diverged = comm_ok ? sh.divg[b] : -1;
            res.residual_sq = fabsf((float)sh.delta[b]); res.tol_sq = sh.tol_sq[b]; res.initial_residual_sq = sh.rsq0[b];
            a.result[b] = res;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
static const int kSmemBudget = 227 * 1024;
// cost of starting a unit, in plane loads (512 x 512 x 64, tails of 6 planes x 7 units against one unit of 64: +5 % instead of -9 %)
#define RING_UNIT_OVERHEAD 2.5

// lines staged per stage = lines_a * TY + lines_b; returns false when the grid lines are too long for a useful ring
static bool ring_config(const DGrid& g, int lines_a, int lines_b, int reserve_bytes, int min_stages, int max_stages,
                        int target_units, RingCfg* out, int consumers = RING_CONSUMERS)
{
    RingCfg c;
    c.split = 0; c.Zm = 0; c.split_t = 0;
    c.consumers = consumers;
    c.pitch = g.cext[0]; c.nx4 = g.cext[0] / 4;
    const int row_bytes = c.pitch * 4;
    // measured on B200 (512^3): TY=4 beats TY=8 for both laplace (6.12 vs 5.85 TB/s) and CG (5.18 vs 5.10 TB/s): smaller
    // stages -> deeper ring; the extra y-halo lines are served by L2
    int ty = g.dim == 3 ? 4 : 16;
    if (g.dim == 3) {
        // measured (tools/sweep_ring.py, 256^3): a CTA wants >= 2048 cells per staged plane - TY=4 146 us, TY=8 114 us, TY=16 114 us
        // per CG iteration, TY=2 538 us; the per-plane mbarrier bookkeeping is amortised over TY * nx cells.  512 -> 4, 256 -> 8, ...
        const int want = 2048 / (g.cext[0] > 0 ? g.cext[0] : 1);
        ty = 1;
        while (ty * 2 <= want && ty < 16) ty *= 2;
        while (ty > 1 && ty / 2 >= g.n[1]) ty /= 2;
    }
    if (const char* e = getenv("PHICUDA_RING_TY")) { const int v = atoi(e); if (v >= 1 && v <= 32) ty = v; }     // tuning knob
    for (;; ty /= 2) {
        if (ty < 1) return false;
        if (ty * c.nx4 > RING_G * consumers) continue;                    // <= RING_G groups per thread
        if (lines_a * ty + lines_b > 128) continue;                            // <= 4 lines per producer lane
        const int stage_bytes = (lines_a * ty + lines_b) * row_bytes;
        const int r = (kSmemBudget - reserve_bytes - 128) / stage_bytes;
        if (const char* e = getenv("PHICUDA_RING_R")) { const int v = atoi(e); if (v >= min_stages && v < max_stages) max_stages = v; }   // tuning knob
        if (r >= min_stages) { c.TY = ty; c.R = r > max_stages ? max_stages : r; c.stage_floats = stage_bytes / 4; break; }
    }
    while (g.dim == 2 && c.TY > 1 && c.TY / 2 >= g.n[1]) c.TY /= 2;
    c.groups = (c.TY * c.nx4 + consumers - 1) / consumers;
    c.shfl_ok = (c.nx4 % 32 == 0) ? 1 : 0;
    { const char* e = getenv("PHICUDA_RING_HINT"); c.hint = (e && e[0] == '1') ? 1 : 0; }
    { const char* e = getenv("PHICUDA_RING_MERGE"); c.merge = (e && e[0] == '0') ? 0 : 1; }
    { const char* e = getenv("PHICUDA_RING_DEBUG"); c.dbg = e ? atoi(e) : 0; }
    if (g.dim == 3) {
        c.nyt = (g.n[1] + c.TY - 1) / c.TY;
        // z chunking: every unit pays two halo planes plus a pipeline start (measured: ~2.5 plane loads, RING_UNIT_OVERHEAD), and
        // the persistent grid of `ctas` CTAs is only fully busy when the unit count is close to a multiple of it -> maximise
        //   utilisation(units, ctas) / (1 + (2 + overhead)/ZC)
        const int ctas = target_units;
        double best = -1.0; int best_nzc = 1;
        const int max_nzc = g.n[2] >= 4 ? g.n[2] / 4 : 1;
        for (int nzc = 1; nzc <= max_nzc; ++nzc) {
            const int zc = (g.n[2] + nzc - 1) / nzc;
            if ((g.n[2] + zc - 1) / zc != nzc) continue;
            const long long units = (long long)c.nyt * nzc * g.batch;
            const long long rounds = (units + ctas - 1) / ctas;
            const double util = (double)units / (double)(rounds * ctas);
            const double score = util / (1.0 + (2.0 + RING_UNIT_OVERHEAD) / zc);
            if (score > best + 1e-9) { best = score; best_nzc = nzc; }
        }
        if (const char* e = getenv("PHICUDA_RING_NZC")) {                     // tuning knob
            const int v = atoi(e);
            if (v >= 1 && v <= g.n[2]) { const int zc = (g.n[2] + v - 1) / v; best_nzc = (g.n[2] + zc - 1) / zc; }
        }
        c.nzc = best_nzc; c.ZC = (g.n[2] + best_nzc - 1) / best_nzc;
        c.units_per_batch = c.nyt * c.nzc;
        c.total_units = c.units_per_batch * g.batch;
    } else {
        c.nyt = (g.n[1] + c.TY - 1) / c.TY; c.nzc = 1; c.ZC = 1;
        c.units_per_batch = c.nyt; c.total_units = c.nyt * g.batch;
    }
    *out = c;
    return true;
}

// every tile and plane of the grid qualifies for the branch-free consumer path
static bool ring_all_fast(const DGrid& g, const DField& f, const RingCfg& c)
{
    if (!c.shfl_ok || c.nx4 * 4 != g.n[0] || c.TY * c.nx4 != c.groups * c.consumers || g.n[1] % c.TY != 0) return false;
    for (int a = 0; a < g.dim; ++a) if (f.klo[a] == PHI_BC_CONST || f.khi[a] == PHI_BC_CONST) return false;
    return c.groups == 1 || c.groups == 2 || c.groups == 4;
}

static void note_ring_launch(int kernel, const RingCfg& c, bool generic, bool dist, bool adaptive, int grid, bool masked = false)
{
    PhiLaunchInfo li; memset(&li, 0, sizeof(li));
    li.kernel = kernel; li.generic = generic; li.dist = dist; li.adaptive = adaptive; li.masked = masked;
    li.TY = c.TY; li.stages = c.R; li.ZC = c.ZC; li.nzc = c.nzc; li.groups = c.groups; li.total_units = c.total_units; li.grid_ctas = grid; li.split = c.split;
    phi_note_launch(li);
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
    if (!ring_config(g, 1, 2, kSmemBudget / 2 + 1024, g.dim == 3 ? 4 : 2, 6, sms * 2, &cfg)) return -100;
    const size_t smem = 128 + (size_t)cfg.R * cfg.stage_floats * 4;
    int grid = sms * 2;
    if (grid > cfg.total_units) grid = cfg.total_units;
    cudaError_t e;
    const bool generic = !ring_all_fast(g, f, cfg);
#define LAUNCH_LAP(D, AX, GEN) do { \
        e = cudaFuncSetAttribute(k_laplace_ring<D, AX, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) k_laplace_ring<D, AX, GEN><<<grid, RING_THREADS, smem, s>>>(g, f, cfg, x, y, coeff); } while (0)
#define LAUNCH_LAP2(D, AX) do { if (generic) LAUNCH_LAP(D, AX, true); else LAUNCH_LAP(D, AX, false); } while (0)
    if (g.dim == 3) { if (axpy) LAUNCH_LAP2(3, true); else LAUNCH_LAP2(3, false); }
    else            { if (axpy) LAUNCH_LAP2(2, true); else LAUNCH_LAP2(2, false); }
#undef LAUNCH_LAP2
#undef LAUNCH_LAP
    if (e != cudaSuccess) return (int)e;
    note_ring_launch(PHI_KERNEL_LAPLACE_RING, cfg, generic, false, false, grid);
    return (int)cudaGetLastError();
}

int phi_launch_cg_ring(const CgLaunch& l, const CommDev* cm, cudaStream_t s)
{
    const DGrid& g = l.g;
    if (g.batch > CG_MAX_BATCH) return -100;
    const bool mask = l.acc != nullptr;
    if (mask && (l.prm.method == PHI_SOLVER_CG_ADAPTIVE || l.prm.matrix_offset != 0.f)) return -100;
    CgRingArgs A;
    const int cgs = (int)((cg_smem_bytes(g.batch) + 127) / 128 * 128);
    const int sms = sm_count();
    // lines staged per stage: pass B (even) = 1 haloed + 3 element-wise arrays = 4 TY + 2; with the obstacle mask as an extra
    // haloed array 5 TY + 4 (pass A: 3 haloed = 3 TY + 6)
    if (!ring_config(g, mask ? 5 : 4, mask ? 6 : 2, cgs, g.dim == 3 ? (mask ? 3 : 4) : 2, RING_MAX_STAGES, sms, &A.cfg)) return -100;
    const int threads = RING_THREADS;
    A.ring_smem_offset = cgs;
    const size_t smem = (size_t)cgs + 128 + (size_t)A.cfg.R * A.cfg.stage_floats * 4;
    int per_sm = 0;
    cudaError_t e;
    const bool generic = mask || !ring_all_fast(g, l.pf, A.cfg);
    const bool dist = cm && cm->n > 1;
    const bool adapt = l.prm.method == PHI_SOLVER_CG_ADAPTIVE;
#define CG_RING_FN2(D, GEN, DI) (adapt ? (const void*)k_cg_ring<D, GEN, DI, true> : (const void*)k_cg_ring<D, GEN, DI, false>)
#define CG_RING_FN(D, GEN) (dist ? CG_RING_FN2(D, GEN, true) : CG_RING_FN2(D, GEN, false))
    const void* fn = g.dim == 3 ? (generic ? CG_RING_FN(3, true) : CG_RING_FN(3, false))
                                : (generic ? CG_RING_FN(2, true) : CG_RING_FN(2, false));
    if (mask) fn = g.dim == 3 ? (dist ? (const void*)k_cg_ring<3, true, true, false, true> : (const void*)k_cg_ring<3, true, false, false, true>)
                              : (dist ? (const void*)k_cg_ring<2, true, true, false, true> : (const void*)k_cg_ring<2, true, false, false, true>);
#undef CG_RING_FN
#undef CG_RING_FN2
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
    if (e != cudaSuccess || per_sm < 1) return -100;
    int grid = sms * per_sm;
    // tail split (see RingCfg): compare the planes the busiest CTA stages per pass, default decomposition vs split
    A.cfg.split = 0; A.cfg.Zm = 0; A.cfg.split_t = 0;
    {
        const char* e = getenv("PHICUDA_RING_SPLIT");                       // tuning knob: 0 = never, 1 = whenever possible
        const int force = e ? atoi(e) : -1;
        const int nyt = A.cfg.nyt, nz = g.n[2];
        if (g.dim == 3 && g.batch == 1 && force != 0 && grid <= CG_MAX_GRID && nyt < grid && nz >= 8) {
            const int spare = grid - nyt, T = (nyt + spare - 1) / spare;
            int best_zm = 0; double best = 1e30;
            for (int zm = 4; zm <= nz - 1; ++zm) {
                const double main_c = zm + 2 + RING_UNIT_OVERHEAD, tail_c = T * (nz - zm + 2 + RING_UNIT_OVERHEAD);
                const double cost = main_c > tail_c ? main_c : tail_c;
                if (cost < best) { best = cost; best_zm = zm; }
            }
            const long long rounds = ((long long)A.cfg.total_units + grid - 1) / grid;
            const double dflt = rounds * (A.cfg.ZC + 2 + RING_UNIT_OVERHEAD);
            if (best_zm > 0 && (force == 1 || best < dflt * 0.97)) {
                A.cfg.split = 1; A.cfg.Zm = best_zm; A.cfg.split_t = T;
                A.cfg.nzc = 2; A.cfg.ZC = best_zm; A.cfg.units_per_batch = A.cfg.total_units = 2 * nyt;
            }
        }
    }
    if (!A.cfg.split && grid > A.cfg.total_units) grid = A.cfg.total_units;
    if (grid > CG_MAX_GRID) grid = CG_MAX_GRID;
    const size_t pf_sb = (size_t)g.cext[0] * g.cext[1] * g.cext[2];
    const size_t arr = ((size_t)pf_sb * g.batch * sizeof(float) + 255) / 256 * 256;
    unsigned char* ws = (unsigned char*)l.workspace;
    CgArgs& a = A.a;
    a.g = g; a.pf = l.pf; a.um = UnitMap();
    a.rhs = l.rhs; a.x = l.x; a.acc = l.acc;
    const size_t hoff = (size_t)g.halo * g.cext[0] * g.cext[1];       // pointers address the first owned plane
    a.r = (float*)ws + hoff; a.d0 = (float*)(ws + arr) + hoff; a.d1 = (float*)(ws + 2 * arr) + hoff;
    a.partials = (double*)(ws + 3 * arr);
    a.result = l.result; a.prm = l.prm;
    if (cm) A.cm = *cm; else { memset(&A.cm, 0, sizeof(A.cm)); A.cm.n = 1; A.cm.lower = A.cm.upper = -1; }
    // measured on 2 GPUs (512 x 512 x 64 slabs): merged 138.0 us / iteration, grid.sync + block-0 send 137.3 us - no gain, so the
    // path that was also verified on 8 GPUs stays the default; PHICUDA_COMM_MERGE=1 selects the merged barrier
    { const char* e = getenv("PHICUDA_COMM_MERGE"); A.comm_merge = (e && e[0] == '1') ? 1 : 0; }
    void* args[] = {&A};
    e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, s);
    if (e != cudaSuccess) { phi_set_error("cg ring: cooperative launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    note_ring_launch(PHI_KERNEL_CG_RING, A.cfg, generic, dist, adapt, grid, mask);
    return 0;
}
