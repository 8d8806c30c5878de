// sgbm_sweep6.cu -- the six row-crossing paths of MODE_HH in one lock-step launch, path-parallel (sm_100a).
//
// Same job and same schedule position as agg_vsweep_kernel (sgbm_agg.cu): between the two horizontal
// scans, the top-down sweep adds its three paths (+1,+1) (0,+1) (-1,+1) to S, the bottom-up sweep writes the sum of its three
// (+1,-1) (0,-1) (-1,-1) to S2.  What changes is the mapping of the work to a warp.  agg_vsweep_kernel gives a warp one
// (column, sweep) and runs the three paths one after the other, 32 lanes x 4 disparities, one warp-wide reduction and two
// shuffles per path and pixel: 147 instructions per (column, row, sweep), of which the per-path overhead (neighbour PRMTs, the
// d = -1 / D sentinels, the minimum) is more than the recurrence itself.  Here a warp owns a COLUMN for BOTH sweeps and
// its eight groups of four lanes own the paths: groups 0-2 the top-down d / dr / dl, groups 4-6 the bottom-up ones (groups 3
// and 7 idle), a lane holds 16 packed registers = 32 disparities in the block layout (SgbmGeom::layout 1), so ONE sgm_step_blk
// advances all six paths of the column: ~75 instructions per (column, row, sweep).
//
// Everything else follows agg_vsweep_kernel: strips of n columns per CTA (all CTAs co-resident), the vertical states in
// registers, the diagonal states through double-buffered shared-memory slots and one mbarrier pair per column (even / odd rows),
// between neighbouring CTAs through a 4-deep ring in global memory whose int16 words carry a phase bit; C (and S) stream through
// a private cp.async ring per warp.  A warp on a CTA boundary steps twice per row: first for the path it hands to the neighbour
// CTA (written at once), then, with the state it receives (loaded at the end of the previous row), for all paths -- the
// hand-over latency overlaps a whole row on both sides, as in agg_vsweep_kernel.
// Block layout, 65..128 disparities (NP = 2) only; everything else keeps agg_vsweep_kernel.
//
// STATUS: opt-in (B2S_SWEEP6=1), bit-exact (tests/test_gpu_sgbm.py::test_six_path_sweep), and SLOWER than agg_vsweep_kernel on a
// B200 at 1080p / 128: 1.62 ms against 0.82 ms (profiles/README.md, r02_sweep6).  The instruction count per (column, row, sweep)
// drops only from 147 to 117 -- the recurrence itself costs four instructions per packed register whatever the mapping, and two
// of the eight lane groups idle -- while each column is now ONE warp whose whole row is a single dependent chain: 13 warps per SM
// instead of 26 to hide it, and the neighbours can only be released after the full six-path step instead of after the two
// diagonal steps.  Measured without any CTA boundary (13 columns, one CTA): 2137 cycles per row against 1004.  Kept as the
// measured answer to "would a fatter warp help the sweep"; the default stays agg_vsweep_kernel.
#include <stdlib.h>

#include <mutex>
#include <type_traits>

#include "sgm_common.cuh"

namespace {

constexpr int V6_R = 8;      // cp.async ring depth (rows)
constexpr int V6_HO = 4;     // slots of a hand-over ring (= HO_SLOTS of sgbm_agg.cu)

struct Vs6Args {
    const int16_t *C;
    int16_t *S, *S2;
    int H, width1, D, P1, P2, n;
    uint32_t *ho; // hand-over rings [2 sweeps][G-1 boundaries][2 dirs][V6_HO][64] u32, memset to 0xFF before the launch
    int *err;
};

// A lane's 16 words (64 bytes) of a pixel row are four 16-byte vectors.  Stored in order, the four lanes of a group would put
// vector i at a stride of 64 bytes and the second group of a quarter-warp 256 bytes further: the same bank group four times.
// So vector i of a lane lives at position i ^ swz inside the lane's 64 bytes, swz = (li >> 1) | (parity << 1), parity = that of the
// path the row belongs to (0 for C / S rows): the 8 lanes of a quarter-warp then cover 8 different bank groups.  vo[i] = byte
// offset of vector i.  The 8-byte readers of the sums undo the permutation (rd_off).
__device__ __forceinline__ void lds16(uint32_t addr, const uint32_t (&vo)[4], uint32_t (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[4 * i]), "=r"(v[4 * i + 1]), "=r"(v[4 * i + 2]), "=r"(v[4 * i + 3]) : "r"(addr + vo[i]) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, const uint32_t (&vo)[4], const uint32_t (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr + vo[i]), "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3]) : "memory");
}
// byte offset inside a 256-byte pixel row of logical 16-byte chunk ci of a row with path parity par
__device__ __forceinline__ uint32_t chunk_off(int ci, int par) { return (uint32_t)(ci ^ ((ci >> 3) | (par << 1))) * 16; }

template <bool PAD>
__global__ void __launch_bounds__(512, 1) agg_vsweep6_kernel(Vs6Args a)
{
    constexpr int CH = 256, DW = 64, N = 16, R = V6_R;
    constexpr int HO_DIR = V6_HO * DW;
    constexpr int STAGEB = 3 * CH; // [C of the top-down row | S of the top-down row | C of the bottom-up row]
    extern __shared__ __align__(16) uint32_t v6_smem[];
    // shared memory: mbarriers [n][2] | slots [2 sweeps][2 parities][2 dirs][n+2][DW] | L buffers [n][2][2 sweeps][3 paths][DW] | rings [n][R][3*DW]
    const int lane = threadIdx.x & 31;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int n = a.n, G = gridDim.x, H = a.H, Dp = 128;
    const int g = lane >> 2, li = lane & 3, sw = g >> 2, q = g & 3; // q: 0 = d (vertical), 1 = dr (from column x-1), 2 = dl (from column x+1), 3 = idle
    const int x = blockIdx.x * n + w;
    const uint32_t BIG = 0x7FFF7FFFu;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(v6_smem);
    const uint32_t MBB = n * 16;
    const uint32_t SLB = 2 * 2 * 2 * (n + 2) * CH;
    const uint32_t LBB = n * 2 * 2 * 3 * CH;
    const uint32_t sl0 = sbase + MBB, lb0 = sl0 + SLB, rg0 = lb0 + LBB;

    uint32_t padmask[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int wd = li * N + i, d0 = (wd >> 3) * 16 + (wd & 7);
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d0 + 8 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    // every slot starts as the out-of-image state; the guard slots (index 0 and n+1) and the slots of idle columns keep it for ever
    for (uint32_t o = threadIdx.x; o < SLB / 4; o += blockDim.x) {
        uint32_t pw = 0;
        if (PAD) { // the word at physical position o % 64 of a slot row of direction `dir` (dr rows: parity 1, dl rows: parity 0)
            const int sdir = (o / ((n + 2) * DW)) & 1, pc = (o & 63) >> 2, ci = pc ^ ((pc >> 3) | ((sdir == 0 ? 1 : 0) << 1));
            const int wd = ci * 4 + (o & 3), d0 = (wd >> 3) * 16 + (wd & 7);
            pw = (d0 >= a.D ? 0x00007FFFu : 0u) | (d0 + 8 >= a.D ? 0x7FFF0000u : 0u);
        }
        v6_smem[MBB / 4 + o] = pw;
    }
    if (threadIdx.x < (unsigned)n * 2) mbar_init(sbase + threadIdx.x * 8, 1);
    __syncthreads();
    if (x >= a.width1) return; // idle warps of the last strip
    const bool first_col = x == 0, last_col = x == a.width1 - 1;
    const bool out_right = w == n - 1 && !last_col, out_left = w == 0 && !first_col; // this column hands a diagonal to another CTA
    const bool has_left = w > 0, has_right = w < n - 1 && !last_col;                  // neighbour columns inside this CTA
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;
    uint32_t ku = li != 0 ? 1u : 0u, kd = li != 3 ? 1u : 0u;
    asm("" : "+r"(ku));
    asm("" : "+r"(kd));
    const uint32_t au = li == 0 ? BIG : 0u, ad = li == 3 ? BIG : 0u;

    const uint32_t PSB = 2 * (n + 2) * CH; // bytes between the two parities of the slots
    const int dir = q == 2 ? 1 : 0;        // slot array of this lane's path (dr: 0, dl: 1)
    uint32_t voq[4], vo0[4]; // vector offsets inside this lane's 64 bytes: rows of this lane's path / C rows
#pragma unroll
    for (int i = 0; i < 4; i++) {
        voq[i] = (uint32_t)(i ^ ((li >> 1) | ((q & 1) << 1))) * 16;
        vo0[i] = (uint32_t)(i ^ (li >> 1)) * 16;
    }
    const uint32_t slq = sl0 + ((sw * 2 * 2 + dir) * (n + 2)) * CH + li * 64;
    const uint32_t in_s = slq + (q == 2 ? w + 2 : w) * CH; // written by column x+1 (dl) / x-1 (dr); guard slots at the image edge
    const uint32_t out_s = slq + (w + 1) * CH;
    const bool diag = q == 1 || q == 2;
    const uint32_t my_mb = sbase + w * 16, mb_left = sbase + (w - 1) * 16, mb_right = sbase + (w + 1) * 16;
    // hand-over rings: boundary b lies between CTA b and b+1; direction 0 crosses it to the right, 1 to the left
    const size_t HJ = (size_t)(G > 1 ? G - 1 : 1) * 2 * HO_DIR;
    const int q_in = out_left ? 1 : 2, q_out = out_left ? 2 : 1; // edge column: the path that arrives from / leaves to the other CTA
    const int b_edge = out_left ? (int)blockIdx.x - 1 : (int)blockIdx.x;
    const uint32_t *in_g = a.ho + sw * HJ + ((size_t)max(b_edge, 0) * 2 + (q_in == 1 ? 0 : 1)) * HO_DIR + li * N;
    uint32_t *out_g = a.ho + sw * HJ + ((size_t)max(b_edge, 0) * 2 + (q_out == 1 ? 0 : 1)) * HO_DIR + li * N;

    // C (and S) stream through a private cp.async ring of R rows per warp: 48 segments of 16 bytes per row
    const uint32_t ring = rg0 + w * (R * STAGEB);
    const long long rowE = (long long)a.width1 * Dp;
    const long long oT = (long long)x * Dp, oB = (long long)(H - 1) * rowE + (long long)x * Dp;
    const int16_t *src0, *src1;
    long long step0, step1;
    {
        const int which0 = lane >> 4, r0 = lane & 15; // segment lane: C or S of the top-down row
        src0 = (which0 ? a.S : a.C) + oT + r0 * 8;
        step0 = rowE;
        src1 = a.C + oB + (lane & 15) * 8;              // segment 32 + lane (lanes 0..15): C of the bottom-up row
        step1 = -rowE;
    }
    // (destination chunks permuted like the readers expect: chunk r of a 256-byte part at r ^ (r >> 3))
    const uint32_t dst0 = ring + (lane >> 4) * CH + chunk_off(lane & 15, 0), dst1 = ring + 2 * CH + chunk_off(lane & 15, 0);
    auto issue = [&](int stage) {
        cp_async16_s(dst0 + stage * STAGEB, src0);
        if (lane < 16) cp_async16_s(dst1 + stage * STAGEB, src1);
        src0 += step0;
        src1 += step1;
    };
#pragma unroll 1
    for (int p = 0; p < R - 1; p++) {
        if (p < H) issue(p);
        cp_async_commit();
    }
    int16_t *spT = a.S + oT + lane * 4, *spB = a.S2 + oB + lane * 4; // this lane's two output words of the row of step t
    const uint32_t cur0 = ring + (sw ? 2 * CH : 0) + li * 64;       // this group's C in a ring stage
    const uint32_t lbw = lb0 + w * (2 * 2 * 3 * CH);                 // this warp's L buffers [2][sweep][path][DW]
    const uint32_t lb_st = lbw + (sw * 3 + q) * CH + li * 64;

    uint32_t T[N], c[N], L[N], Tpre[N];
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = Tpre[i] = padmask[i];
    cp_async_wait<R - 2>();
    __syncwarp();
    lds16(cur0, vo0, c);
    const uint32_t rd0 = chunk_off(lane >> 1, 0) + (lane & 1) * 8, rd1 = chunk_off(lane >> 1, 1) + (lane & 1) * 8; // the sums' 8-byte reads

    int stage = 1 % R, pstage = R - 1;
    uint32_t pin = PSB, pout = 0; // parity offsets of the slots read (step t-1) and written (step t)
    auto ho_poll = [&](int t) { // Tpre holds the words loaded at the end of step t-1: poll until every int16 carries phase (t-1)
        const uint32_t phase = (((t - 1) >> 2) & 1) ? 0x80008000u : 0u;
        const uint32_t *p = in_g + ((t - 1) & (V6_HO - 1)) * DW;
        int spins = 0;
        unsigned long long t0 = 0;
        while (true) {
            uint32_t bad = 0;
            if (q == q_in) {
#pragma unroll
                for (int i = 0; i < N; i++) bad |= (Tpre[i] ^ phase) & 0x80008000u;
            }
            if (__all_sync(0xffffffffu, bad == 0)) break;
            if (wait_expired(++spins, t0, a.err)) {
                *(volatile int *)a.err = 1;
                break;
            }
            if (q == q_in) {
#pragma unroll
                for (int i = 0; i < N; i += 4)
                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Tpre[i]), "=r"(Tpre[i + 1]), "=r"(Tpre[i + 2]), "=r"(Tpre[i + 3]) : "l"(p + i) : "memory");
            }
        }
    };
    auto rows = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
#pragma unroll 1
        for (int t = 0; t < H; t++) {
            // a column signals even rows on its first mbarrier and odd rows on the second (a neighbour may run one row ahead);
            // at t = 0 the wait is for the phase before the first one, which counts as complete
            const uint32_t mb_off = ((t - 1) & 1) * 8, par_in = ((t - 1) >> 1) & 1;
            if (EDGE) {
                // ---- first the path that leaves the CTA: its state comes from the neighbour column inside the CTA ----
                if (out_left ? has_right : has_left) mbar_wait(out_left ? mb_right + mb_off : mb_left + mb_off, par_in, a.err);
                uint32_t T2[N], L2[N];
#pragma unroll
                for (int i = 0; i < N; i++) T2[i] = T[i];
                if (q == q_out) lds16(in_s + pin, voq, T2);
                sgm_step_blk<N, 4, PAD>(T2, c, L2, padmask, P1v, P2mP1v, ku, au, kd, ad);
                if (q == q_out) {
                    const uint32_t phase = ((t >> 2) & 1) ? 0x80008000u : 0u;
                    uint32_t *p = out_g + (t & (V6_HO - 1)) * DW;
#pragma unroll
                    for (int i = 0; i < N; i += 4)
                        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"((T2[i] & 0x7FFF7FFFu) | phase),
                                     "r"((T2[i + 1] & 0x7FFF7FFFu) | phase), "r"((T2[i + 2] & 0x7FFF7FFFu) | phase), "r"((T2[i + 3] & 0x7FFF7FFFu) | phase)
                                     : "memory");
                }
                // ---- then all paths, with the state that arrived from the other CTA (loaded at the end of the previous row) ----
                if (t > 0) ho_poll(t);
                if (q == q_in) {
#pragma unroll
                    for (int i = 0; i < N; i++) T[i] = t > 0 ? (Tpre[i] & 0x7FFF7FFFu) : padmask[i];
                } else if (q == q_out)
                    lds16(in_s + pin, voq, T);
            } else {
                if (has_left) mbar_wait(mb_left + mb_off, par_in, a.err);
                if (has_right) mbar_wait(mb_right + mb_off, par_in, a.err);
                if (diag) lds16(in_s + pin, voq, T);
            }
            sgm_step_blk<N, 4, PAD>(T, c, L, padmask, P1v, P2mP1v, ku, au, kd, ad);
            if (EDGE ? q == q_in : diag) sts16(out_s + pout, voq, T); // (an edge column's outgoing path is already in the global ring)
            const uint32_t lb = (t & 1) * (2 * 3 * CH);
            if (q < 3) sts16(lb_st + lb, voq, L);
            __syncwarp();
            if (lane == 0) mbar_arrive(my_mb + (t & 1) * 8); // this column's row-t states are in their slots
            if (EDGE && q == q_in) { // the neighbour CTA wrote the state of the next row early in ITS row t: fetch it now
                const uint32_t *p = in_g + (t & (V6_HO - 1)) * DW;
#pragma unroll
                for (int i = 0; i < N; i += 4)
                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Tpre[i]), "=r"(Tpre[i + 1]), "=r"(Tpre[i + 2]), "=r"(Tpre[i + 3]) : "l"(p + i) : "memory");
            }
            // ---- sums: lane l owns words 2l, 2l+1 of both sweeps' outputs ----
            {
                const uint32_t la = lbw + lb;
                uint32_t s0[2], s1[2], s2[2], sv[2];
                lds_s<2>(la + rd0, s0);          // path d  (parity 0)
                lds_s<2>(la + CH + rd1, s1);     // path dr (parity 1)
                lds_s<2>(la + 2 * CH + rd0, s2); // path dl (parity 0)
                lds_s<2>(ring + ((stage + R - 1) % R) * STAGEB + CH + rd0, sv); // S of the top-down row (the stage of step t)
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    uint32_t v = __viaddmin_u16x2(s0[i], s1[i], BIG);
                    v = __viaddmin_u16x2(v, s2[i], BIG);
                    sv[i] = __viaddmin_u16x2(sv[i], v, BIG); // saturating sums (L >= 0: any order)
                }
                stcg_regs<2>(spT, sv);
                lds_s<2>(la + 3 * CH + rd0, s0);
                lds_s<2>(la + 4 * CH + rd1, s1);
                lds_s<2>(la + 5 * CH + rd0, s2);
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    uint32_t v = __viaddmin_u16x2(s0[i], s1[i], BIG);
                    sv[i] = __viaddmin_u16x2(v, s2[i], BIG);
                }
                stcg_regs<2>(spB, sv);
            }
            spT += rowE;
            spB -= rowE;
            __syncwarp();
            if (t + R - 1 < H) issue(pstage);
            cp_async_commit();
            cp_async_wait<R - 2>(); // row t+1 has landed
            __syncwarp();
            lds16(cur0 + stage * STAGEB, vo0, c);
            pstage = pstage + 1 == R ? 0 : pstage + 1;
            stage = stage + 1 == R ? 0 : stage + 1;
            pin = pout;
            pout ^= PSB;
        }
    };
    if (out_left || out_right) rows(std::true_type{});
    else rows(std::false_type{});
}

} // namespace

// columns per CTA for the six-path sweep, 0 = not applicable (the caller falls back to agg_vsweep_kernel)
int vsweep6_cols(const b2s_ctx *c, const SgbmGeom &g)
{
    const char *on = getenv("B2S_SWEEP6");
    if (!on || atoi(on) == 0 || getenv("B2S_AGG_LEGACY")) return 0; // opt-in: slower than agg_vsweep_kernel (see the header)
    if (g.mode != 1 || g.D <= 64 || g.D > 128) return 0;
    int n = (g.width1 + c->num_sms - 1) / c->num_sms;
    if (n < 8) n = g.width1 < 8 ? g.width1 : 8;
    if (const char *e = getenv("B2S_VSWEEP_COLS")) { // test hook: force narrow strips so that small images span several CTAs
        const int v = atoi(e);
        if (v >= 2 && v <= 16 && (g.width1 + v - 1) / v <= c->num_sms) n = v;
    }
    return (n >= 2 && n <= 16) ? n : 0; // (16 warps per CTA: the kernel needs ~120 registers per thread)
}

cudaError_t launch_vsweep6(b2s_ctx *c, int n, bool cooperative)
{
    const SgbmGeom &g = c->g;
    if (g.layout != 1 || g.NP != 2 || n < 2 || n > 16) return cudaErrorInvalidValue;
    if (cudaError_t te = wait_timeout_init()) return te;
    const int G = (g.width1 + n - 1) / n;
    Vs6Args a;
    a.C = c->C.as<int16_t>();
    a.S = c->S.as<int16_t>();
    a.S2 = c->S2.as<int16_t>();
    a.H = g.H; a.width1 = g.width1; a.D = g.D; a.P1 = g.P1; a.P2 = g.P2; a.n = n;
    const size_t ho_bytes = (size_t)2 * (G > 1 ? G - 1 : 1) * 2 * V6_HO * 64 * sizeof(uint32_t);
    cudaError_t e = c->agg_ho.ensure(ho_bytes);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->agg_ho.p, 0xFF, ho_bytes, c->stream)) != cudaSuccess) return e; // phase 1 everywhere (steps 0..3 write phase 0)
    a.ho = c->agg_ho.as<uint32_t>();
    a.err = c->agg_err;
    const size_t smem = (size_t)n * 16 + (size_t)2 * 2 * 2 * (n + 2) * 256 + (size_t)n * 2 * 2 * 3 * 256 + (size_t)n * V6_R * 3 * 256;
    const bool pad = g.D != g.Dp;
    auto go = [&](auto kern, int slot) -> cudaError_t {
        static std::once_flag once[64][2];
        cudaError_t ee = cudaSuccess;
        std::call_once(once[c->device & 63][slot], [&] { ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); });
        if (ee != cudaSuccess) return ee;
        int occ = 0;
        if ((ee = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, n * 32, smem)) != cudaSuccess) return ee;
        if (occ < 1 || G > occ * c->num_sms) return cudaErrorCooperativeLaunchTooLarge; // all strips must be co-resident
        if (!cooperative) {
            kern<<<G, n * 32, smem, c->stream>>>(a);
        } else { // (see launch_vsweep_t in sgbm_agg.cu)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(n * 32);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = c->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if ((ee = cudaLaunchKernelEx(&cfg, kern, a)) != cudaSuccess) return ee;
        }
        c->launches++;
        return cudaGetLastError();
    };
    return pad ? go(agg_vsweep6_kernel<true>, 0) : go(agg_vsweep6_kernel<false>, 1);
}
